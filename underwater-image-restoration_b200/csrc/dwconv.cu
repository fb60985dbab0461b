// Depthwise 3x3 convolution on token (NHWC) tensors with the two GELUs of LeFF / FRFN fused in
// (LeFF: AST.py:299-301,312-321; FRFN gate: AST.py:334-336,360-367).
//
//   h1 = gelu(u[..., :Ch]) ; v = dwconv3x3(h1) + bias ; h2 = gelu(v) [* gelu(u[..., Ch:2Ch])]
//
// CTA = 16x16 pixels x 32 channels, lane <-> channel (128 B coalesced per pixel), halo tile staged
// once in shared memory after the GELU so erf is evaluated once per element.  HBM-bound:
// fwd reads u (x1.27 halo, mostly L2 hits) and writes v,h2; bwd reads dh2,v,u and writes du.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int TS = 16;          // tile side
constexpr int HS = TS + 2;      // halo side
constexpr int CG = 32;          // channels per CTA
constexpr int DW_THREADS = 256; // 8 warps, 2 tile rows each

__global__ void __launch_bounds__(DW_THREADS) dwconv_fwd_kernel(const float* __restrict__ u, long long ld_u,
                                                                const float* __restrict__ weight,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ v, float* __restrict__ h2,
                                                                int H, int W, int Ch, int mode, int tiles_x) {
    __shared__ float h1s[HS * HS][CG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * CG + lane;
    const bool cok = c < Ch;
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
    const float* ub = u + (long long)b * H * W * ld_u;

    for (int pix = warp; pix < HS * HS; pix += DW_THREADS / 32) {
        const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
        float val = 0.f;
        if (cok && y >= 0 && y < H && x >= 0 && x < W) val = gelu_f(ub[((long long)y * W + x) * ld_u + c]);
        h1s[pix][lane] = val;
    }
    float wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wgt[k] = cok ? weight[c * 9 + k] : 0.f;
    const float bv = cok ? bias[c] : 0.f;
    __syncthreads();
    if (!cok) return;

#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int ly = warp * 2 + rr;  // local row 0..15
        const int y = ty0 + ly;
        if (y >= H) break;
        float win[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            win[ky][1] = h1s[(ly + ky) * HS + 0][lane];
            win[ky][2] = h1s[(ly + ky) * HS + 1][lane];
        }
#pragma unroll
        for (int lx = 0; lx < TS; ++lx) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                win[ky][0] = win[ky][1];
                win[ky][1] = win[ky][2];
                win[ky][2] = h1s[(ly + ky) * HS + lx + 2][lane];
            }
            const int x = tx0 + lx;
            if (x < W) {
                float acc = bv;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(win[ky][kx], wgt[ky * 3 + kx], acc);
                const long long tok = ((long long)b * H + y) * W + x;
                if (v) v[tok * Ch + c] = acc;
                float o = gelu_f(acc);
                if (mode == 1) o *= gelu_f(u[tok * ld_u + Ch + c]);
                h2[tok * Ch + c] = o;
            }
        }
    }
}

// persistent over tiles: grid = (P, Ch/32); each CTA accumulates dweight/dbias partials in registers
__global__ void __launch_bounds__(DW_THREADS) dwconv_bwd_kernel(const float* __restrict__ dh2,
                                                                const float* __restrict__ u, long long ld_u,
                                                                const float* __restrict__ v,
                                                                const float* __restrict__ weight,
                                                                float* __restrict__ du, float* __restrict__ partials,
                                                                int B, int H, int W, int Ch, int mode, int tiles_x,
                                                                int tiles_per_img) {
    extern __shared__ __align__(16) float smem[];
    float(*dvs)[CG] = reinterpret_cast<float(*)[CG]>(smem);
    float(*h1s)[CG] = reinterpret_cast<float(*)[CG]>(smem + HS * HS * CG);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * CG + lane;
    const bool cok = c < Ch;

    float wgt[9], dwt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        wgt[k] = cok ? weight[c * 9 + k] : 0.f;
        dwt[k] = 0.f;
    }
    float dbs = 0.f;

    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * TS, tx0 = (tl % tiles_x) * TS;
        const long long base = (long long)b * H * W;
        __syncthreads();
        for (int pix = warp; pix < HS * HS; pix += DW_THREADS / 32) {
            const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
            float dvv = 0.f, h1 = 0.f;
            if (cok && y >= 0 && y < H && x >= 0 && x < W) {
                const long long tok = base + (long long)y * W + x;
                const float vv = v[tok * Ch + c];
                float d = dh2[tok * Ch + c] * gelu_grad_f(vv);
                if (mode == 1) d *= gelu_f(u[tok * ld_u + Ch + c]);
                dvv = d;
                h1 = gelu_f(u[tok * ld_u + c]);
            }
            dvs[pix][lane] = dvv;
            h1s[pix][lane] = h1;
        }
        __syncthreads();
        if (cok) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int ly = warp * 2 + rr;
                const int y = ty0 + ly;
                if (y >= H) break;
                for (int lx = 0; lx < TS; ++lx) {
                    const int x = tx0 + lx;
                    if (x >= W) break;
                    // dh1[y,x] = sum_k dv[y+1-ky, x+1-kx] w[ky,kx]; local halo index of (y,x) is (ly+1, lx+1)
                    float dh1 = 0.f;
                    const float dvc = dvs[(ly + 1) * HS + lx + 1][lane];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            dh1 = fmaf(dvs[(ly + 2 - ky) * HS + (lx + 2 - kx)][lane], wgt[ky * 3 + kx], dh1);
                            dwt[ky * 3 + kx] = fmaf(dvc, h1s[(ly + ky) * HS + lx + kx][lane], dwt[ky * 3 + kx]);
                        }
                    dbs += dvc;
                    const long long tok = base + (long long)y * W + x;
                    const float uc = u[tok * ld_u + c];
                    du[tok * ld_u + c] = dh1 * gelu_grad_f(uc);
                    if (mode == 1) {
                        const float u2 = u[tok * ld_u + Ch + c];
                        du[tok * ld_u + Ch + c] = dh2[tok * Ch + c] * gelu_f(v[tok * Ch + c]) * gelu_grad_f(u2);
                    }
                }
            }
        }
    }
    // reduce the 10 per-channel partial sums across the 8 warps
    __syncthreads();
    float* red = smem;  // [8][10][32]
#pragma unroll
    for (int k = 0; k < 9; ++k) red[(warp * 10 + k) * CG + lane] = dwt[k];
    red[(warp * 10 + 9) * CG + lane] = dbs;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 10 * CG; idx += DW_THREADS) {
        const int k = idx / CG, l = idx % CG;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < DW_THREADS / 32; ++w) s += red[(w * 10 + k) * CG + l];
        const int cc = blockIdx.y * CG + l;
        if (cc < Ch) partials[((long long)blockIdx.x * 10 + k) * Ch + cc] = s;
    }
}

__global__ void dwconv_param_reduce_kernel(const float* __restrict__ partials, float* __restrict__ dweight,
                                           float* __restrict__ dbias, int P, int Ch) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 10 * Ch) return;
    const int k = idx / Ch, c = idx % Ch;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += partials[((long long)p * 10 + k) * Ch + c];
    if (k < 9) dweight[c * 9 + k] = s;
    else dbias[c] = s;
}

int bwd_ctas(int B, int H, int W, int Ch) {
    const int tiles = B * uwr_cdiv(H, TS) * uwr_cdiv(W, TS);
    const int groups = uwr_cdiv(Ch, CG);
    int p = uwr_cdiv(2 * uwr_sm_count(), groups);
    if (p > tiles) p = tiles;
    return p < 1 ? 1 : p;
}

}  // namespace

extern "C" int uwr_dwconv_gelu_fwd(const float* u, long long ld_u, const float* weight, const float* bias, float* v,
                                   float* h2, int B, int H, int W, int Ch, int mode, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(u && weight && bias && h2, "uwr_dwconv_gelu_fwd: null pointer");
    UWR_REQUIRE(mode == 0 || mode == 1, "uwr_dwconv_gelu_fwd: mode must be 0 (LeFF) or 1 (FRFN gate)");
    UWR_REQUIRE(ld_u >= (mode == 1 ? 2 * Ch : Ch), "uwr_dwconv_gelu_fwd: ld_u too small");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_dwconv_gelu_fwd: bad batch %d", B);
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    dim3 grid(tx * ty, uwr_cdiv(Ch, CG), B);
    dwconv_fwd_kernel<<<grid, DW_THREADS, 0, stream>>>(u, ld_u, weight, bias, v, h2, H, W, Ch, mode, tx);
    UWR_CHECK_LAUNCH("dwconv_fwd_kernel");
    return 0;
}

extern "C" size_t uwr_dwconv_gelu_bwd_workspace_bytes(int B, int H, int W, int Ch) {
    return (size_t)bwd_ctas(B, H, W, Ch) * 10 * (size_t)Ch * sizeof(float);
}

extern "C" int uwr_dwconv_gelu_bwd(const float* dh2, const float* u, long long ld_u, const float* v,
                                   const float* weight, float* du, float* dweight, float* dbias, float* workspace,
                                   int B, int H, int W, int Ch, int mode, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dh2 && u && v && weight && du && dweight && dbias && workspace, "uwr_dwconv_gelu_bwd: null pointer");
    UWR_REQUIRE(mode == 0 || mode == 1, "uwr_dwconv_gelu_bwd: mode must be 0 or 1");
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    const int P = bwd_ctas(B, H, W, Ch);
    constexpr int smem_bytes = 2 * HS * HS * CG * (int)sizeof(float);
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(dwconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured = true;
    }
    dim3 grid(P, uwr_cdiv(Ch, CG));
    dwconv_bwd_kernel<<<grid, DW_THREADS, smem_bytes, stream>>>(dh2, u, ld_u, v, weight, du, workspace, B, H, W, Ch,
                                                               mode, tx, tx * ty);
    UWR_CHECK_LAUNCH("dwconv_bwd_kernel");
    dwconv_param_reduce_kernel<<<uwr_cdiv(10 * Ch, 128), 128, 0, stream>>>(workspace, dweight, dbias, P, Ch);
    UWR_CHECK_LAUNCH("dwconv_param_reduce_kernel");
    return 0;
}
