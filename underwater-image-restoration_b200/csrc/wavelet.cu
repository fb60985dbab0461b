// Haar "wavelet" mode of the src/model family (use_dwt = "Wavelet": EncoderBlock model.py:64-88, FDFP block.py:532-552)
// on token (NHWC) tensors.  The reference's DWT_2D / IDWT_2D (wave_modules.py:9-181) are NOT a per-channel wavelet
// transform: their (C/4, C, 2, 2) filters are one 2x2 Haar tap pattern broadcast over every (out, in) channel pair, so
//
//   DWT  : out[b, n*C/4 + c', y, x]      = sum_{dy,dx} w_n[dy][dx] * S[b, 2y+dy, 2x+dx],  S = sum over ALL input channels
//   IDWT : out[b, 4g + o, 2y+dy, 2x+dx]  = w_o[dy][dx] * T[b, g, y, x],                   T = sum of channels 4g .. 4g+3
//
// (n, o in {ll, lh, hl, hh}; w = Haar taps 0.5 * (+-1)), and their hand-written autograd backward passes
// (DWT_function.backward, IDWT_function.backward) are not the adjoints of those maps.  Both are reproduced here as
// the reference computes them — including the raw NCHW-memory reshapes of IDWT_function.backward — because "results
// identical to the reference" includes its gradients:
//
//   DWT bwd : din[b, any c, 2y+dy, 2x+dx]  = sum_n w_n[dy][dx] * R_n[b, y, x],
//             R_n = sum_{j in block n of the (c n)-reordered channels} dout[b, (j % 4) * C/4 + j / 4, y, x]
//   IDWT bwd: din[b, n*C/4 + c', Y, X]     = V_n[b, (c'*h*w + Y*w + X) mod (h*w/16)],
//             V_n[b, y, x] = sum_{ch < 16C} sum_{dy,dx} w_n[dy][dx] * flat_b[ch*(h*w/4) + (2y+dy)*(w/2) + 2x+dx],
//             flat_b = the NCHW-contiguous memory of dout[b] (C, 2h, 2w) (wave_modules.py:86-114).
//
// Tensors here are tokens: (B, H*W, C) row-major.  Tiny bandwidth passes (one read + one write of the tensor).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

// Haar tap sign of sub-band n at (dy, dx), magnitude 0.5: ll ++++, lh (rows differ), hl (columns differ), hh
__device__ __forceinline__ float haar_w(int n, int dy, int dx) {
    const int sy = (n == 1 || n == 3) ? (dy ? -1 : 1) : 1;   // lh, hh: w[i][j] carries hi[i] (i = dy)
    const int sx = (n == 2 || n == 3) ? (dx ? -1 : 1) : 1;   // hl, hh: w[i][j] carries hi[j] (j = dx)
    return 0.5f * (float)(sy * sx);
}

// ---- DWT forward: in (B, 2h*2w, C) -> out (B, h*w, C).  One warp per output pixel.
__global__ void __launch_bounds__(256) dwt_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int h,
                                                      int w, int C) {
    uwr_pdl_enter();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp; pix < (long long)B * h * w; pix += nwarps) {
        const int x = (int)(pix % w), y = (int)((pix / w) % h);
        const long long b = pix / ((long long)w * h);
        float S[2][2];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float* src = in + ((b * 2 * h + 2 * y + dy) * (2LL * w) + 2 * x + dx) * C;
                float s = 0.f;
                for (int c = lane; c < C; c += 32) s += src[c];
                S[dy][dx] = warp_sum(s);
            }
        float sub[4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
            sub[n] = haar_w(n, 0, 0) * S[0][0] + haar_w(n, 0, 1) * S[0][1] + haar_w(n, 1, 0) * S[1][0] + haar_w(n, 1, 1) * S[1][1];
        float* dst = out + pix * C;
        const int q = C / 4;
        for (int c = lane; c < C; c += 32) dst[c] = sub[c / q];
    }
}

// ---- DWT backward (reference formula): dout (B, h*w, C) -> din (B, 2h*2w, C)
__global__ void __launch_bounds__(256) dwt_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B, int h,
                                                      int w, int C) {
    uwr_pdl_enter();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int q = C / 4;
    for (long long pix = warp; pix < (long long)B * h * w; pix += nwarps) {
        const int x = (int)(pix % w), y = (int)((pix / w) % h);
        const long long b = pix / ((long long)w * h);
        const float* src = dout + pix * C;
        // R_n = sum over j in [n*q, (n+1)*q) of dout channel (j % 4) * q + j / 4
        float R[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = lane; j < C; j += 32) {
            const float v = src[(j & 3) * q + (j >> 2)];
            const int n = j / q;
            R[0] += n == 0 ? v : 0.f; R[1] += n == 1 ? v : 0.f; R[2] += n == 2 ? v : 0.f; R[3] += n == 3 ? v : 0.f;
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) R[n] = warp_sum(R[n]);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float v = haar_w(0, dy, dx) * R[0] + haar_w(1, dy, dx) * R[1] + haar_w(2, dy, dx) * R[2] + haar_w(3, dy, dx) * R[3];
                float* dst = din + ((b * 2 * h + 2 * y + dy) * (2LL * w) + 2 * x + dx) * C;
                for (int c = lane; c < C; c += 32) dst[c] = v;
            }
    }
}

// ---- IDWT forward: in (B, h*w, C) -> out (B, 2h*2w, C); thread per (coarse pixel, group of 4 channels)
__global__ void __launch_bounds__(256) idwt_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int h,
                                                       int w, int C) {
    uwr_pdl_enter();
    const int G = C / 4;
    const long long total = (long long)B * h * w * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const long long pix = i / G;
        const int x = (int)(pix % w), y = (int)((pix / w) % h);
        const long long b = pix / ((long long)w * h);
        const float4 v = *reinterpret_cast<const float4*>(in + pix * C + 4 * g);
        const float T = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float* dst = out + ((b * 2 * h + 2 * y + dy) * (2LL * w) + 2 * x + dx) * C + 4 * g;
                *reinterpret_cast<float4*>(dst) = make_float4(haar_w(0, dy, dx) * T, haar_w(1, dy, dx) * T,
                                                              haar_w(2, dy, dx) * T, haar_w(3, dy, dx) * T);
            }
    }
}

// ---- IDWT backward (reference formula), step 1: V (B, 4, h*w/16) from dout tokens (B, 2h*2w, C)
// flat NCHW index f of dout[b] -> channel f / (4hw), row (f % 4hw) / 2w, column f % 2w.  One warp per (b, y, x).
__global__ void __launch_bounds__(256) idwt_bwd_v_kernel(const float* __restrict__ dout, float* __restrict__ V, int B, int h,
                                                         int w, int C) {
    uwr_pdl_enter();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int h4 = h / 4, w4 = w / 4;
    const long long plane = 4LL * h * w;          // elements of one (2h x 2w) output channel plane
    const long long rplane = (long long)h * w / 4;  // elements of one (h/2 x w/2) plane of the reshaped tensor
    for (long long it = warp; it < (long long)B * h4 * w4; it += nwarps) {
        const int x = (int)(it % w4), y = (int)((it / w4) % h4);
        const long long b = it / ((long long)w4 * h4);
        const float* base = dout + b * plane * C;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = lane; ch < 16 * C; ch += 32) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long long f = ch * rplane + (long long)(2 * y + dy) * (w / 2) + 2 * x + dx;
                    const long long c = f / plane, rem = f - c * plane;
                    const float v = base[rem * C + c];      // token layout: (Y * 2w + X) * C + c, and rem = Y * 2w + X
#pragma unroll
                    for (int n = 0; n < 4; ++n) acc[n] = fmaf(haar_w(n, dy, dx), v, acc[n]);
                }
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const float s = warp_sum(acc[n]);
            if (lane == 0) V[(b * 4 + n) * ((long long)h4 * w4) + (long long)y * w4 + x] = s;
        }
    }
}

// step 2: din (B, h*w, C): channel n*C/4 + c' at (Y, X) <- V_n[(c'*h*w + Y*w + X) mod (h*w/16)]
__global__ void __launch_bounds__(256) idwt_bwd_scatter_kernel(const float* __restrict__ V, float* __restrict__ din, int B,
                                                               int h, int w, int C) {
    uwr_pdl_enter();
    const long long total = (long long)B * h * w * C;
    const int q = C / 4;
    const long long m = (long long)h * w / 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long yx = pix % ((long long)h * w);
        const long long b = pix / ((long long)h * w);
        const int n = c / q, cp = c - n * q;
        const long long pos = ((long long)cp * h * w + yx) % m;
        din[i] = V[(b * 4 + n) * m + pos];
    }
}

int wl_blocks(long long work) {
    long long b = (work + 255) / 256;
    const long long cap = 8LL * uwr_sm_count();
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// h, w: the COARSE grid (the fine grid is 2h x 2w).  C % 4 == 0.
extern "C" int uwr_haar_dwt_fwd(const float* in, float* out, int B, int h, int w, int C, uwr_stream_t stream_) {
    UWR_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && C % 4 == 0, "uwr_haar_dwt_fwd: bad args (C %% 4 == 0)");
    (void)uwr_launch_pdl(dwt_fwd_kernel, dim3(wl_blocks((long long)B * h * w * 32)), dim3(256), 0, (cudaStream_t)stream_, in, out, B, h, w, C);
    UWR_CHECK_LAUNCH("dwt_fwd_kernel");
    return 0;
}

extern "C" int uwr_haar_dwt_bwd(const float* dout, float* din, int B, int h, int w, int C, uwr_stream_t stream_) {
    UWR_REQUIRE(dout && din && B > 0 && h > 0 && w > 0 && C % 4 == 0, "uwr_haar_dwt_bwd: bad args (C %% 4 == 0)");
    (void)uwr_launch_pdl(dwt_bwd_kernel, dim3(wl_blocks((long long)B * h * w * 32)), dim3(256), 0, (cudaStream_t)stream_, dout, din, B, h, w, C);
    UWR_CHECK_LAUNCH("dwt_bwd_kernel");
    return 0;
}

extern "C" int uwr_haar_idwt_fwd(const float* in, float* out, int B, int h, int w, int C, uwr_stream_t stream_) {
    UWR_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && C % 4 == 0, "uwr_haar_idwt_fwd: bad args (C %% 4 == 0)");
    (void)uwr_launch_pdl(idwt_fwd_kernel, dim3(wl_blocks((long long)B * h * w * (C / 4))), dim3(256), 0, (cudaStream_t)stream_, in, out, B, h, w, C);
    UWR_CHECK_LAUNCH("idwt_fwd_kernel");
    return 0;
}

// workspace >= B * 4 * (h*w/16) floats; the reference's reshapes need h and w divisible by 4
extern "C" int uwr_haar_idwt_bwd(const float* dout, float* din, float* workspace, int B, int h, int w, int C,
                                 uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dout && din && workspace && B > 0 && C % 4 == 0, "uwr_haar_idwt_bwd: bad args (C %% 4 == 0)");
    UWR_REQUIRE(h % 4 == 0 && w % 4 == 0 && h > 0 && w > 0, "uwr_haar_idwt_bwd: h and w must be multiples of 4");
    (void)uwr_launch_pdl(idwt_bwd_v_kernel, dim3(wl_blocks((long long)B * (h / 4) * (w / 4) * 32)), dim3(256), 0, stream, dout, workspace, B, h, w, C);
    UWR_CHECK_LAUNCH("idwt_bwd_v_kernel");
    (void)uwr_launch_pdl(idwt_bwd_scatter_kernel, dim3(wl_blocks((long long)B * h * w * C)), dim3(256), 0, stream, workspace, din, B, h, w, C);
    UWR_CHECK_LAUNCH("idwt_bwd_scatter_kernel");
    return 0;
}
