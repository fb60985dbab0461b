// Elementwise passes of the FFT amplitude / phase up-sampler (SpectralTransformer.UpSample.forward,
// src/Models/SpectralTransformer.py:174-188) and their backward, one fused kernel per reference chain:
//
//   polar_split : mag = abs(f), pha = angle(f)                       (lines 176-177)
//   polar_join  : z = mag cos(pha) + i mag sin(pha)                  (lines 181-183: real / imag / complex)
//   cabs        : |ifft2(.)|                                          (line 186)
//   leaky_relu  : the LeakyReLU(0.1) between the two 1x1 convs of amp_fuse / pha_fuse (lines 166-169)
//   even_scatter: post(|.|) lives on the even pixels of the 2H x 2W grid, post(0) = bias elsewhere
//                 (the (2,2)-tiled spectrum identity)
//
// Complex tensors are interleaved (re, im) fp32 pairs (torch.view_as_real layout).  Every kernel is a
// grid-stride pass with 128-bit accesses: algorithmic bytes = the tensors named in the signature, HBM-bound.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int EW_THREADS = 256;

inline int ew_grid(long long work) {
    long long b = (work + EW_THREADS - 1) / EW_THREADS;
    const long long cap = 8LL * uwr_sm_count();
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// n complex numbers, two per thread iteration (one float4 in, one float2 to each output)
__global__ void polar_split_fwd_kernel(const float4* __restrict__ f, float2* __restrict__ mag, float2* __restrict__ pha,
                                       long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = f[i];
        mag[i] = make_float2(hypotf(v.x, v.y), hypotf(v.z, v.w));
        pha[i] = make_float2(atan2f(v.y, v.x), atan2f(v.w, v.z));
    }
}

__device__ __forceinline__ float2 polar_split_grad(float re, float im, float dm, float dp) {
    const float r2 = re * re + im * im;
    if (r2 == 0.f) return make_float2(0.f, 0.f);   // torch: abs'(0) = 0, angle'(0) = 0
    const float ir = rsqrtf(r2), ir2 = 1.0f / r2;
    return make_float2(dm * re * ir - dp * im * ir2, dm * im * ir + dp * re * ir2);
}

__global__ void polar_split_bwd_kernel(const float4* __restrict__ f, const float2* __restrict__ dmag,
                                       const float2* __restrict__ dpha, float4* __restrict__ df, long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = f[i];
        const float2 dm = dmag[i], dp = dpha[i];
        const float2 a = polar_split_grad(v.x, v.y, dm.x, dp.x), b = polar_split_grad(v.z, v.w, dm.y, dp.y);
        df[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}

__global__ void polar_join_fwd_kernel(const float2* __restrict__ mag, const float2* __restrict__ pha,
                                      float4* __restrict__ z, long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float2 m = mag[i], p = pha[i];
        float s0, c0, s1, c1;
        sincosf(p.x, &s0, &c0);
        sincosf(p.y, &s1, &c1);
        z[i] = make_float4(m.x * c0, m.x * s0, m.y * c1, m.y * s1);
    }
}

__global__ void polar_join_bwd_kernel(const float2* __restrict__ mag, const float2* __restrict__ pha,
                                      const float4* __restrict__ dz, float2* __restrict__ dmag, float2* __restrict__ dpha,
                                      long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float2 m = mag[i], p = pha[i];
        const float4 d = dz[i];
        float s0, c0, s1, c1;
        sincosf(p.x, &s0, &c0);
        sincosf(p.y, &s1, &c1);
        dmag[i] = make_float2(d.x * c0 + d.y * s0, d.z * c1 + d.w * s1);
        dpha[i] = make_float2(m.x * (d.y * c0 - d.x * s0), m.y * (d.w * c1 - d.z * s1));
    }
}

__global__ void cabs_fwd_kernel(const float4* __restrict__ z, float2* __restrict__ a, long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = z[i];
        a[i] = make_float2(hypotf(v.x, v.y), hypotf(v.z, v.w));
    }
}

__global__ void cabs_bwd_kernel(const float4* __restrict__ z, const float2* __restrict__ da, float4* __restrict__ dz,
                                long long n2) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = z[i];
        const float2 d = da[i];
        const float r0 = v.x * v.x + v.y * v.y, r1 = v.z * v.z + v.w * v.w;
        const float i0 = r0 > 0.f ? d.x * rsqrtf(r0) : 0.f, i1 = r1 > 0.f ? d.y * rsqrtf(r1) : 0.f;
        dz[i] = make_float4(v.x * i0, v.y * i0, v.z * i1, v.w * i1);
    }
}

// y = x > 0 ? x : slope * x ; backward selects on the sign of the saved OUTPUT (slope > 0 keeps the sign)
__global__ void leaky_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4, float slope, int rnd) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = x[i];
        v.x = v.x > 0.f ? v.x : slope * v.x; v.y = v.y > 0.f ? v.y : slope * v.y;
        v.z = v.z > 0.f ? v.z : slope * v.z; v.w = v.w > 0.f ? v.w : slope * v.w;
        if (rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        y[i] = v;
    }
}

__global__ void leaky_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ dy, float4* __restrict__ dx,
                                 long long n4, float slope) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = y[i], d = dy[i];
        dx[i] = make_float4(v.x > 0.f ? d.x : slope * d.x, v.y > 0.f ? d.y : slope * d.y, v.z > 0.f ? d.z : slope * d.z,
                            v.w > 0.f ? d.w : slope * d.w);
    }
}

// exact (erf) GELU as nn.GELU(): FDFP between its two 1x1 convs (block.py:541-546), Mlp.act (AST.py:285-291)
__global__ void gelu_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4, int rnd) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        float4 o = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
        if (rnd) o = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
        y[i] = o;
    }
}

__global__ void gelu_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, float4* __restrict__ dx,
                                long long n4) {
    uwr_pdl_enter();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[i], d = dy[i];
        dx[i] = make_float4(d.x * gelu_grad_f(v.x), d.y * gelu_grad_f(v.y), d.z * gelu_grad_f(v.z), d.w * gelu_grad_f(v.w));
    }
}

// out (B, 2H, 2W, C): even pixels <- y (B, H, W, C), every other pixel <- bias.  One float4 of out per thread step.
__global__ void even_scatter_kernel(const float* __restrict__ y, const float* __restrict__ bias, float* __restrict__ out,
                                    int H, int W, int C, long long total4) {
    uwr_pdl_enter();
    const int c4n = C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const long long pix = i / c4n;
        const int x = (int)(pix % (2 * W));
        const long long t = pix / (2 * W);
        const int yy = (int)(t % (2 * H));
        const long long b = t / (2 * H);
        float4 v;
        if (((x | yy) & 1) == 0) v = *reinterpret_cast<const float4*>(y + (((b * H + (yy >> 1)) * W + (x >> 1)) * C + c4 * 4));
        else v = *reinterpret_cast<const float4*>(bias + c4 * 4);
        *reinterpret_cast<float4*>(out + i * 4) = v;
    }
}

// dy (B, H, W, C) <- the even pixels of dout (B, 2H, 2W, C)
__global__ void even_gather_kernel(const float* __restrict__ dout, float* __restrict__ dy, int H, int W, int C,
                                   long long total4) {
    uwr_pdl_enter();
    const int c4n = C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        const long long pix = i / c4n;
        const int x = (int)(pix % W);
        const long long t = pix / W;
        const int yy = (int)(t % H);
        const long long b = t / H;
        *reinterpret_cast<float4*>(dy + i * 4) =
            *reinterpret_cast<const float4*>(dout + (((b * 2 * H + 2 * yy) * (2LL * W) + 2 * x) * C + c4 * 4));
    }
}

}  // namespace

#define EW_CHECK_N2(who)                                                                     \
    UWR_REQUIRE(n > 0 && n % 2 == 0, who ": element count must be a positive multiple of 2")

extern "C" int uwr_polar_split_fwd(const float* f, float* mag, float* pha, long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(f && mag && pha, "uwr_polar_split_fwd: null pointer");
    EW_CHECK_N2("uwr_polar_split_fwd");
    (void)uwr_launch_pdl(polar_split_fwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float4*)f, (float2*)mag, (float2*)pha, n / 2);
    UWR_CHECK_LAUNCH("polar_split_fwd_kernel");
    return 0;
}

extern "C" int uwr_polar_split_bwd(const float* f, const float* dmag, const float* dpha, float* df, long long n,
                                   uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(f && dmag && dpha && df, "uwr_polar_split_bwd: null pointer");
    EW_CHECK_N2("uwr_polar_split_bwd");
    (void)uwr_launch_pdl(polar_split_bwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float4*)f, (const float2*)dmag,
                                                                     (const float2*)dpha, (float4*)df, n / 2);
    UWR_CHECK_LAUNCH("polar_split_bwd_kernel");
    return 0;
}

extern "C" int uwr_polar_join_fwd(const float* mag, const float* pha, float* z, long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(mag && pha && z, "uwr_polar_join_fwd: null pointer");
    EW_CHECK_N2("uwr_polar_join_fwd");
    (void)uwr_launch_pdl(polar_join_fwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float2*)mag, (const float2*)pha, (float4*)z, n / 2);
    UWR_CHECK_LAUNCH("polar_join_fwd_kernel");
    return 0;
}

extern "C" int uwr_polar_join_bwd(const float* mag, const float* pha, const float* dz, float* dmag, float* dpha,
                                  long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(mag && pha && dz && dmag && dpha, "uwr_polar_join_bwd: null pointer");
    EW_CHECK_N2("uwr_polar_join_bwd");
    (void)uwr_launch_pdl(polar_join_bwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float2*)mag, (const float2*)pha,
                                                                    (const float4*)dz, (float2*)dmag, (float2*)dpha, n / 2);
    UWR_CHECK_LAUNCH("polar_join_bwd_kernel");
    return 0;
}

extern "C" int uwr_cabs_fwd(const float* z, float* a, long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(z && a, "uwr_cabs_fwd: null pointer");
    EW_CHECK_N2("uwr_cabs_fwd");
    (void)uwr_launch_pdl(cabs_fwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float4*)z, (float2*)a, n / 2);
    UWR_CHECK_LAUNCH("cabs_fwd_kernel");
    return 0;
}

extern "C" int uwr_cabs_bwd(const float* z, const float* da, float* dz, long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(z && da && dz, "uwr_cabs_bwd: null pointer");
    EW_CHECK_N2("uwr_cabs_bwd");
    (void)uwr_launch_pdl(cabs_bwd_kernel, dim3(ew_grid(n / 2)), dim3(EW_THREADS), 0, stream, (const float4*)z, (const float2*)da, (float4*)dz, n / 2);
    UWR_CHECK_LAUNCH("cabs_bwd_kernel");
    return 0;
}

extern "C" int uwr_gelu_fwd(const float* x, float* y, long long n, int round_out, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && y && n > 0 && n % 4 == 0, "uwr_gelu_fwd: n %% 4 == 0 required");
    (void)uwr_launch_pdl(gelu_fwd_kernel, dim3(ew_grid(n / 4)), dim3(EW_THREADS), 0, stream, (const float4*)x, (float4*)y, n / 4,
                                                              round_out && uwr_round_outputs());
    UWR_CHECK_LAUNCH("gelu_fwd_kernel");
    return 0;
}

extern "C" int uwr_gelu_bwd(const float* x, const float* dy, float* dx, long long n, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && dy && dx && n > 0 && n % 4 == 0, "uwr_gelu_bwd: n %% 4 == 0 required");
    (void)uwr_launch_pdl(gelu_bwd_kernel, dim3(ew_grid(n / 4)), dim3(EW_THREADS), 0, stream, (const float4*)x, (const float4*)dy, (float4*)dx, n / 4);
    UWR_CHECK_LAUNCH("gelu_bwd_kernel");
    return 0;
}

extern "C" int uwr_leaky_relu_fwd(const float* x, float* y, long long n, float slope, int round_out, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && y && n > 0 && n % 4 == 0 && slope > 0.f, "uwr_leaky_relu_fwd: n %% 4 == 0 and slope > 0 required");
    (void)uwr_launch_pdl(leaky_fwd_kernel, dim3(ew_grid(n / 4)), dim3(EW_THREADS), 0, stream, (const float4*)x, (float4*)y, n / 4, slope,
                                                               round_out && uwr_round_outputs());
    UWR_CHECK_LAUNCH("leaky_fwd_kernel");
    return 0;
}

extern "C" int uwr_leaky_relu_bwd(const float* y, const float* dy, float* dx, long long n, float slope,
                                  uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(y && dy && dx && n > 0 && n % 4 == 0 && slope > 0.f, "uwr_leaky_relu_bwd: n %% 4 == 0 and slope > 0 required");
    (void)uwr_launch_pdl(leaky_bwd_kernel, dim3(ew_grid(n / 4)), dim3(EW_THREADS), 0, stream, (const float4*)y, (const float4*)dy, (float4*)dx, n / 4, slope);
    UWR_CHECK_LAUNCH("leaky_bwd_kernel");
    return 0;
}

extern "C" int uwr_even_scatter(const float* y, const float* bias, float* out, int B, int H, int W, int C,
                                uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(y && bias && out && C % 4 == 0 && B > 0 && H > 0 && W > 0, "uwr_even_scatter: bad args (C %% 4 == 0)");
    const long long total4 = (long long)B * 4 * H * W * (C / 4);
    (void)uwr_launch_pdl(even_scatter_kernel, dim3(ew_grid(total4)), dim3(EW_THREADS), 0, stream, y, bias, out, H, W, C, total4);
    UWR_CHECK_LAUNCH("even_scatter_kernel");
    return 0;
}

extern "C" int uwr_even_gather(const float* dout, float* dy, int B, int H, int W, int C, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dout && dy && C % 4 == 0 && B > 0 && H > 0 && W > 0, "uwr_even_gather: bad args (C %% 4 == 0)");
    const long long total4 = (long long)B * H * W * (C / 4);
    (void)uwr_launch_pdl(even_gather_kernel, dim3(ew_grid(total4)), dim3(EW_THREADS), 0, stream, dout, dy, H, W, C, total4);
    UWR_CHECK_LAUNCH("even_gather_kernel");
    return 0;
}
