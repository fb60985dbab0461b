// Pixel losses of LossFunction.getloss (src/Losses/losses.py:54-81,182-213; luminanceLoss.py:5-21):
// value and gradient in one pass over pred/truth, deterministic two-level reduction.
// HBM-bound: 3 * B*C*H*W * 4 bytes (read pred, truth; write grad).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 1024;

// kind: 0 L1, 1 L1withColor, 2 charbonnier, 3 L2, 4 mean((clamp01(p) - clamp01(t))^2) (torchPSNR, no gradient)
__global__ void __launch_bounds__(LOSS_THREADS) pixel_loss_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ truth,
                                                                  float* __restrict__ grad,
                                                                  float* __restrict__ partials, int kind, int B,
                                                                  int C, long long HW, float inv_n, float inv_pix,
                                                                  float inv_div) {
    uwr_pdl_enter();
    __shared__ float red[3][LOSS_THREADS / 32];
    float s_abs = 0.f, s_sq = 0.f, s_lum = 0.f;
    const long long npix = (long long)B * HW;
    const float ycoef[3] = {0.299f, 0.587f, 0.114f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / HW, pix = i % HW;
        const long long base = b * C * HW + pix;
        if (kind == 1) {
            float d[3], lum = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c] = pred[base + c * HW] - truth[base + c * HW];
                lum += ycoef[c] * d[c];
                s_abs += fabsf(d[c]);
                s_sq += d[c] * d[c];
            }
            s_lum += lum * lum;
            if (grad) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float sg = d[c] > 0.f ? 1.f : (d[c] < 0.f ? -1.f : 0.f);
                    grad[base + c * HW] =
                        (0.5f * 2.f * d[c] * inv_n + 0.25f * sg * inv_n + 0.25f * 2.f * lum * ycoef[c] * inv_pix) * inv_div;
                }
            }
        } else {
            for (int c = 0; c < C; ++c) {
                float d = pred[base + c * HW] - truth[base + c * HW];
                if (kind == 4)  // torchPSNR clamps both images to [0, 1] first (ModelTrainer.py:17-21)
                    d = fminf(fmaxf(pred[base + c * HW], 0.f), 1.f) - fminf(fmaxf(truth[base + c * HW], 0.f), 1.f);
                float gv;
                if (kind == 0) {
                    s_abs += fabsf(d);
                    gv = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n * inv_div;
                } else if (kind == 2) {
                    const float r = sqrtf(d * d + 1e-6f);
                    s_abs += r;
                    gv = d / r * inv_n;
                } else {
                    s_sq += d * d;
                    gv = 2.f * d * inv_n * inv_div;
                }
                if (grad) grad[base + c * HW] = gv;
            }
        }
    }
    s_abs = warp_sum(s_abs);
    s_sq = warp_sum(s_sq);
    s_lum = warp_sum(s_lum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[0][warp] = s_abs;
        red[1][warp] = s_sq;
        red[2][warp] = s_lum;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += red[threadIdx.x][w];
        partials[blockIdx.x * 3 + threadIdx.x] = s;
    }
}

__global__ void pixel_loss_final_kernel(const float* __restrict__ partials, float* __restrict__ out, int nblocks,
                                        int kind, float inv_n, float inv_pix, float inv_div) {
    uwr_pdl_enter();
    __shared__ double sh[3][32];
    const int lane = threadIdx.x;  // 32 threads
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = lane; i < nblocks; i += 32) {
        a += partials[i * 3 + 0];
        b += partials[i * 3 + 1];
        c += partials[i * 3 + 2];
    }
    sh[0][lane] = a;
    sh[1][lane] = b;
    sh[2][lane] = c;
    __syncthreads();
    if (lane == 0) {
        double A = 0, Bq = 0, Cl = 0;
        for (int i = 0; i < 32; ++i) {
            A += sh[0][i];
            Bq += sh[1][i];
            Cl += sh[2][i];
        }
        double loss;
        if (kind == 0) loss = A * inv_n * inv_div;
        else if (kind == 1) loss = (0.5 * Bq * inv_n + 0.25 * A * inv_n + 0.25 * Cl * inv_pix) * inv_div;
        else if (kind == 2) loss = A * inv_n;
        else if (kind == 4) loss = Bq * inv_n;
        else loss = Bq * inv_n * inv_div;
        out[0] = (float)loss;
    }
}

}  // namespace

extern "C" int uwr_pixel_loss(const float* pred, const float* truth, float* out, float* grad, float* workspace,
                              int kind, int B, int C, int H, int W, int batch_divisor, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(pred && truth && out && workspace, "uwr_pixel_loss: null pointer");
    UWR_REQUIRE(kind >= 0 && kind <= 4, "uwr_pixel_loss: kind %d unsupported", kind);
    UWR_REQUIRE(kind != 4 || grad == nullptr, "uwr_pixel_loss: kind 4 (clamped MSE) has no gradient");
    UWR_REQUIRE(kind != 1 || C == 3, "uwr_pixel_loss: L1withColor needs 3 channels");
    UWR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && batch_divisor > 0, "uwr_pixel_loss: bad shape");
    const long long HW = (long long)H * W;
    const long long npix = (long long)B * HW;
    int blocks = (int)((npix + LOSS_THREADS - 1) / LOSS_THREADS);
    const int cap = 4 * uwr_sm_count() < LOSS_MAX_BLOCKS ? 4 * uwr_sm_count() : LOSS_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    const float inv_n = (float)(1.0 / ((double)npix * C));
    const float inv_pix = (float)(1.0 / (double)npix);
    const float inv_div = (float)(1.0 / ((double)batch_divisor * C));
    (void)uwr_launch_pdl(pixel_loss_kernel, dim3(blocks), dim3(LOSS_THREADS), 0, stream, pred, truth, grad, workspace, kind, B, C, HW, inv_n, inv_pix,
                                                          inv_div);
    UWR_CHECK_LAUNCH("pixel_loss_kernel");
    (void)uwr_launch_pdl(pixel_loss_final_kernel, dim3(1), dim3(32), 0, stream, workspace, out, blocks, kind, inv_n, inv_pix, inv_div);
    UWR_CHECK_LAUNCH("pixel_loss_final_kernel");
    return 0;
}
