// Device versions of two "fflMix" terms (src/Losses/losses.py:108-117) and of the SSIM validation metric:
//
//   * Gradient_Loss (losses.py:162-181): L1 between the 3x3 Laplacians (valid convolution) of pred and truth,
//     value + gradient w.r.t. pred;
//   * pytorch_msssim (third party, restated in SURVEY.md Appendix C): separable 11-tap Gaussian (sigma 1.5), valid
//     convolution, the per-scale cs / ssim maps, their per-plane means (forward) and the gradient w.r.t. the first
//     image through the blur's adjoint (backward).  MS_SSIM's 5-scale product-of-powers is (planes x 5) scalar algebra
//     and stays on the host side (uwr/ssim.py); ssim() of one scale is ModelTrainer.torchSSIM (ModelTrainer.py:23-24).
//
// Images are NCHW fp32 planes (B*C planes of H x W).  All kernels are tile-based stencil passes through shared
// memory; algorithmic bytes: Laplacian 3 planes (read p, t; write grad), SSIM forward 2 reads + 3 map writes per
// pixel, backward 3 map reads + 2 reads + 1 write -> HBM / latency bound (12.6 MB per tensor at B = 16, 256x256).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int ST = 16;            // output tile side
constexpr int WIN = 11;           // Gaussian taps
constexpr int SH = ST + WIN - 1;  // staged tile side (26)
constexpr int SS_THREADS = 256;

struct Gauss {
    float g[WIN];
};

// ------------------------------------------------------------------------------------------- Laplacian L1
// pass 1: s = sign(lap(p - t)) on the valid region (0 elsewhere), per-block partial sums of |lap|
__global__ void __launch_bounds__(256) lap_sign_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                       float* __restrict__ s, float* __restrict__ partials, int planes,
                                                       int H, int W) {
    uwr_pdl_enter();
    __shared__ float red[8];
    const long long total = (long long)planes * H * W;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        float sv = 0.f;
        if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
            const float l = (p[i - W] - t[i - W]) + (p[i + W] - t[i + W]) + (p[i - 1] - t[i - 1]) + (p[i + 1] - t[i + 1]) -
                            4.f * (p[i] - t[i]);
            acc += fabsf(l);
            sv = l > 0.f ? 1.f : (l < 0.f ? -1.f : 0.f);
        }
        s[i] = sv;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += red[w];
        partials[blockIdx.x] = a;
    }
}

// pass 2: grad = lap^T(s) / n (the kernel is symmetric; s is zero outside the valid region), loss = sum / n
__global__ void __launch_bounds__(256) lap_grad_kernel(const float* __restrict__ s, const float* __restrict__ partials,
                                                       int nparts, float* __restrict__ grad, float* __restrict__ loss,
                                                       int planes, int H, int W, float inv_n) {
    uwr_pdl_enter();
    const long long total = (long long)planes * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (grad == nullptr) break;
        const int x = (int)(i % W), y = (int)((i / W) % H);
        float g = -4.f * s[i];
        if (y > 0) g += s[i - W];
        if (y < H - 1) g += s[i + W];
        if (x > 0) g += s[i - 1];
        if (x < W - 1) g += s[i + 1];
        grad[i] = g * inv_n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double a = 0.0;
        for (int k = 0; k < nparts; ++k) a += partials[k];
        loss[0] = (float)(a * inv_n);
    }
}

// ------------------------------------------------------------------------------------------------ SSIM
// forward of one scale.  grid = (tiles_x, tiles_y, planes).  Outputs per valid pixel the derivatives of the
// per-pixel value (cs, or ssim = l * cs when `full`) w.r.t. the blurred quantities a = blur(X), e = blur(X^2),
// h = blur(XY) (maps[0..2], skipped when maps == NULL), and per-block partial sums of cs and of ssim.
__global__ void __launch_bounds__(SS_THREADS)
ssim_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ maps,
                float* __restrict__ part_cs, float* __restrict__ part_ss, int H, int W, Gauss gw, float C1, float C2,
                int full) {
    uwr_pdl_enter();
    __shared__ float xs[SH * SH], ys[SH * SH];
    __shared__ float hz[5][SH * ST];
    __shared__ float red[2][SS_THREADS / 32];
    const int Ho = H - WIN + 1, Wo = W - WIN + 1;
    const int plane = blockIdx.z, x0 = blockIdx.x * ST, y0 = blockIdx.y * ST;
    const float* Xp = X + (long long)plane * H * W;
    const float* Yp = Y + (long long)plane * H * W;
    for (int i = threadIdx.x; i < SH * SH; i += SS_THREADS) {
        const int r = i / SH, c = i - r * SH, yy = y0 + r, xx = x0 + c;
        const bool ok = yy < H && xx < W;
        xs[i] = ok ? Xp[(long long)yy * W + xx] : 0.f;
        ys[i] = ok ? Yp[(long long)yy * W + xx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SH * ST; i += SS_THREADS) {   // horizontal pass: 26 rows x 16 columns
        const int r = i / ST, c = i - r * ST;
        float a = 0.f, b = 0.f, e = 0.f, f = 0.f, h = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float xv = xs[r * SH + c + k], yv = ys[r * SH + c + k], w = gw.g[k];
            a = fmaf(w, xv, a); b = fmaf(w, yv, b);
            e = fmaf(w, xv * xv, e); f = fmaf(w, yv * yv, f); h = fmaf(w, xv * yv, h);
        }
        hz[0][i] = a; hz[1][i] = b; hz[2][i] = e; hz[3][i] = f; hz[4][i] = h;
    }
    __syncthreads();
    const int ty = threadIdx.x / ST, tx = threadIdx.x - ty * ST;
    const int oy = y0 + ty, ox = x0 + tx;
    float cs = 0.f, ss = 0.f;
    if (oy < Ho && ox < Wo) {
        float a = 0.f, b = 0.f, e = 0.f, f = 0.f, h = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const int j = (ty + k) * ST + tx;
            const float w = gw.g[k];
            a = fmaf(w, hz[0][j], a); b = fmaf(w, hz[1][j], b); e = fmaf(w, hz[2][j], e);
            f = fmaf(w, hz[3][j], f); h = fmaf(w, hz[4][j], h);
        }
        const float A1 = 2.f * a * b + C1, A2 = 2.f * (h - a * b) + C2;
        const float B1 = a * a + b * b + C1, B2 = (e - a * a) + (f - b * b) + C2;
        cs = A2 / B2;
        ss = (A1 / B1) * cs;
        if (maps != nullptr) {
            float ga, ge, gh;
            const float iB2 = 1.f / B2;
            if (full) {   // value = A1 A2 / (B1 B2)
                const float iB1 = 1.f / B1, den = iB1 * iB2;
                ga = ((2.f * b * A2 - 2.f * b * A1) - A1 * A2 * (2.f * a * iB1 - 2.f * a * iB2)) * den;
                ge = -A1 * A2 * den * iB2;
                gh = 2.f * A1 * den;
            } else {      // value = A2 / B2
                ga = (-2.f * b + 2.f * a * A2 * iB2) * iB2;
                ge = -A2 * iB2 * iB2;
                gh = 2.f * iB2;
            }
            const long long o = ((long long)plane * 3) * Ho * Wo + (long long)oy * Wo + ox;
            maps[o] = ga;
            maps[o + (long long)Ho * Wo] = ge;
            maps[o + 2LL * Ho * Wo] = gh;
        }
    }
    cs = warp_sum(cs);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cs; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < SS_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
        const long long o = (long long)plane * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x;
        part_cs[o] = a;
        part_ss[o] = b;
    }
}

// means[plane] = sum of that plane's tile partials / (Ho*Wo), for cs and ssim
__global__ void ssim_mean_kernel(const float* __restrict__ part_cs, const float* __restrict__ part_ss, int tiles,
                                 float inv_count, float* __restrict__ mean_cs, float* __restrict__ mean_ss) {
    uwr_pdl_enter();
    const int plane = blockIdx.x;
    __shared__ double sh[2][32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < tiles; i += 32) {
        a += part_cs[(long long)plane * tiles + i];
        b += part_ss[(long long)plane * tiles + i];
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < 32; ++i) { s0 += sh[0][i]; s1 += sh[1][i]; }
        mean_cs[plane] = (float)(s0 * inv_count);
        mean_ss[plane] = (float)(s1 * inv_count);
    }
}

// backward of one scale: dX(q) = coef[plane] * sum_t g[t] g[u] (Ga + 2 X(q) Ge + Y(q) Gh)(q - (t, u))  [+ 0.25 * dXc(q / 2)]
// (adjoint of the valid blur; the maps are zero outside [0,Ho) x [0,Wo)); dXc = gradient of the next coarser scale
// (avg_pool2d(2) adjoint), NULL for the coarsest.
__global__ void __launch_bounds__(SS_THREADS)
ssim_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ maps,
                const float* __restrict__ coef, const float* __restrict__ dXc, float* __restrict__ dX, int H, int W,
                Gauss gw) {
    uwr_pdl_enter();
    __shared__ float ms[3][SH * SH];
    __shared__ float hz[3][SH * ST];
    const int Ho = H - WIN + 1, Wo = W - WIN + 1;
    const int plane = blockIdx.z, x0 = blockIdx.x * ST, y0 = blockIdx.y * ST;
    const float* mp = maps + ((long long)plane * 3) * Ho * Wo;
    // staged map tile covers output coordinates (y0 - 10 .. y0 + 15, x0 - 10 .. x0 + 15)
    for (int i = threadIdx.x; i < SH * SH; i += SS_THREADS) {
        const int r = i / SH, c = i - r * SH, yy = y0 - (WIN - 1) + r, xx = x0 - (WIN - 1) + c;
        const bool ok = yy >= 0 && yy < Ho && xx >= 0 && xx < Wo;
        const long long o = (long long)yy * Wo + xx;
        ms[0][i] = ok ? mp[o] : 0.f;
        ms[1][i] = ok ? mp[o + (long long)Ho * Wo] : 0.f;
        ms[2][i] = ok ? mp[o + 2LL * Ho * Wo] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SH * ST; i += SS_THREADS) {   // horizontal: sum_u g[u] M(r, c + 10 - u)
        const int r = i / ST, c = i - r * ST;
        float a = 0.f, e = 0.f, h = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const int j = r * SH + c + (WIN - 1) - k;
            const float w = gw.g[k];
            a = fmaf(w, ms[0][j], a); e = fmaf(w, ms[1][j], e); h = fmaf(w, ms[2][j], h);
        }
        hz[0][i] = a; hz[1][i] = e; hz[2][i] = h;
    }
    __syncthreads();
    const int ty = threadIdx.x / ST, tx = threadIdx.x - ty * ST;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) {
        float a = 0.f, e = 0.f, h = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const int j = (ty + (WIN - 1) - k) * ST + tx;
            const float w = gw.g[k];
            a = fmaf(w, hz[0][j], a); e = fmaf(w, hz[1][j], e); h = fmaf(w, hz[2][j], h);
        }
        const long long o = (long long)plane * H * W + (long long)y * W + x;
        float g = coef[plane] * (a + 2.f * X[o] * e + Y[o] * h);
        if (dXc != nullptr) g += 0.25f * dXc[(long long)plane * (H / 2) * (W / 2) + (long long)(y >> 1) * (W / 2) + (x >> 1)];
        dX[o] = g;
    }
}

// avg_pool2d(kernel 2, stride 2) of even-sized planes, both images in one launch
__global__ void __launch_bounds__(256) avgpool2_pair_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                            float* __restrict__ Xo, float* __restrict__ Yo, int planes,
                                                            int H, int W) {
    uwr_pdl_enter();
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)planes * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
        const long long pl = i / ((long long)Wo * Ho);
        const long long o = pl * H * W + (long long)(2 * y) * W + 2 * x;
        Xo[i] = 0.25f * (X[o] + X[o + 1] + X[o + W] + X[o + W + 1]);
        Yo[i] = 0.25f * (Y[o] + Y[o + 1] + Y[o + W] + Y[o + W + 1]);
    }
}

Gauss make_gauss() {   // pytorch_msssim._fspecial_gauss_1d(11, 1.5)
    Gauss g;
    double s = 0.0, v[WIN];
    for (int i = 0; i < WIN; ++i) {
        const double c = i - WIN / 2;
        v[i] = exp(-(c * c) / (2.0 * 1.5 * 1.5));
        s += v[i];
    }
    for (int i = 0; i < WIN; ++i) g.g[i] = (float)(v[i] / s);
    return g;
}

int ew_blocks2(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 8LL * uwr_sm_count();
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" size_t uwr_laplacian_l1_workspace_bytes(int planes, int H, int W) {
    return ((size_t)planes * H * W + 8192) * sizeof(float);
}

extern "C" int uwr_laplacian_l1_loss(const float* pred, const float* truth, float* loss, float* grad, float* workspace,
                                     int planes, int H, int W, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(pred && truth && loss && workspace && planes > 0 && H >= 3 && W >= 3, "uwr_laplacian_l1_loss: bad args");
    const long long total = (long long)planes * H * W;
    int blocks = ew_blocks2(total);
    if (blocks > 8192) blocks = 8192;
    float* s = workspace + 8192;
    (void)uwr_launch_pdl(lap_sign_kernel, dim3(blocks), dim3(256), 0, stream, pred, truth, s, workspace, planes, H, W);
    UWR_CHECK_LAUNCH("lap_sign_kernel");
    const float inv_n = (float)(1.0 / ((double)planes * (H - 2) * (W - 2)));
    (void)uwr_launch_pdl(lap_grad_kernel, dim3(grad ? blocks : 1), dim3(256), 0, stream, s, workspace, blocks, grad, loss, planes, H, W, inv_n);
    UWR_CHECK_LAUNCH("lap_grad_kernel");
    return 0;
}

extern "C" size_t uwr_ssim_workspace_bytes(int planes, int H, int W) {
    const size_t tiles = (size_t)uwr_cdiv(H - WIN + 1, ST) * uwr_cdiv(W - WIN + 1, ST);
    return 2 * (size_t)planes * tiles * sizeof(float);
}

// one scale: mean_cs[planes], mean_ss[planes]; maps (planes, 3, H-10, W-10) or NULL; full = 1: maps differentiate ssim
extern "C" int uwr_ssim_scale_fwd(const float* X, const float* Y, float* maps, float* mean_cs, float* mean_ss,
                                  float* workspace, int planes, int H, int W, float data_range, int full,
                                  uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(X && Y && mean_cs && mean_ss && workspace && planes > 0 && planes <= 65535, "uwr_ssim_scale_fwd: bad args");
    UWR_REQUIRE(H >= WIN && W >= WIN, "uwr_ssim_scale_fwd: planes must be at least 11 x 11 (got %d x %d)", H, W);
    const int Ho = H - WIN + 1, Wo = W - WIN + 1;
    dim3 grid(uwr_cdiv(Wo, ST), uwr_cdiv(Ho, ST), planes);
    const int tiles = grid.x * grid.y;
    float* pcs = workspace;
    float* pss = workspace + (size_t)planes * tiles;
    const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
    (void)uwr_launch_pdl(ssim_fwd_kernel, dim3(grid), dim3(SS_THREADS), 0, stream, X, Y, maps, pcs, pss, H, W, make_gauss(), C1, C2, full);
    UWR_CHECK_LAUNCH("ssim_fwd_kernel");
    (void)uwr_launch_pdl(ssim_mean_kernel, dim3(planes), dim3(32), 0, stream, pcs, pss, tiles, (float)(1.0 / ((double)Ho * Wo)), mean_cs, mean_ss);
    UWR_CHECK_LAUNCH("ssim_mean_kernel");
    return 0;
}

// dX = coef[plane] * d(sum over the plane of the per-pixel value)/dX  [+ avg_pool adjoint of dX_coarse]
extern "C" int uwr_ssim_scale_bwd(const float* X, const float* Y, const float* maps, const float* coef,
                                  const float* dX_coarse, float* dX, int planes, int H, int W, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(X && Y && maps && coef && dX && planes > 0 && planes <= 65535, "uwr_ssim_scale_bwd: bad args");
    UWR_REQUIRE(H >= WIN && W >= WIN && (dX_coarse == nullptr || (H % 2 == 0 && W % 2 == 0)),
                "uwr_ssim_scale_bwd: planes >= 11 x 11, even sides below a coarser scale");
    dim3 grid(uwr_cdiv(W, ST), uwr_cdiv(H, ST), planes);
    (void)uwr_launch_pdl(ssim_bwd_kernel, dim3(grid), dim3(SS_THREADS), 0, stream, X, Y, maps, coef, dX_coarse, dX, H, W, make_gauss());
    UWR_CHECK_LAUNCH("ssim_bwd_kernel");
    return 0;
}

extern "C" int uwr_avgpool2_pair(const float* X, const float* Y, float* Xo, float* Yo, int planes, int H, int W,
                                 uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(X && Y && Xo && Yo && planes > 0 && H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2,
                "uwr_avgpool2_pair: even plane sides required");
    (void)uwr_launch_pdl(avgpool2_pair_kernel, dim3(ew_blocks2((long long)planes * (H / 2) * (W / 2))), dim3(256), 0, stream, X, Y, Xo, Yo, planes, H, W);
    UWR_CHECK_LAUNCH("avgpool2_pair_kernel");
    return 0;
}
