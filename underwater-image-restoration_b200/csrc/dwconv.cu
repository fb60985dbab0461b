// Depthwise 3x3 convolution on token (NHWC) tensors with the two GELUs of LeFF / FRFN fused in
// (LeFF: AST.py:299-301,312-321; FRFN gate: AST.py:334-336,360-367).
//
//   h1 = gelu(u[..., :Ch]) ; v = dwconv3x3(h1) + bias ; h2 = gelu(v) [* gelu(u[..., Ch:2Ch])]
//
// CTA = 16x16 pixels x 32 channels, lane <-> channel (128 B coalesced per pixel).  The halo tile is
// staged once in shared memory (after the GELU in the forward, so erf is evaluated once per
// element); global loads are issued in batches of 8 per warp to keep enough bytes in flight.
// Backward consumes dv = dL/dv (produced by the linear2 data-gradient GEMM epilogue,
// UWR_EPI_MUL_DGELU, or by uwr_gelu_gate_bwd) so only ONE tensor needs a halo:
//   dh1[p] = sum_k dv[p+1-k] w[k] ; du = dh1 * gelu'(u) ; dw[k] = sum_p gelu(u[p]) dv[p+1-k]
// HBM-bound: fwd reads u (x1.27 halo, mostly L2 hits) and writes v,h2; bwd reads dv (x1.27), u and
// writes du.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int TS = 16;          // tile side
constexpr int HS = TS + 2;      // halo side
constexpr int CG = 32;          // channels per CTA
constexpr int DW_THREADS = 256; // 8 warps, 2 tile rows each
constexpr int NWARP = DW_THREADS / 32;
constexpr int BATCH = 8;        // global loads in flight per warp while staging

__global__ void __launch_bounds__(DW_THREADS) dwconv_fwd_kernel(const float* __restrict__ u, long long ld_u,
                                                                const float* __restrict__ weight,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ v, float* __restrict__ h2,
                                                                int H, int W, int Ch, int mode, int tiles_x, int rnd,
                                                                int v_is_dgelu) {
    __shared__ float h1s[HS * HS][CG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * CG + lane;
    const bool cok = c < Ch;
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
    const float* ub = u + (long long)b * H * W * ld_u;

    for (int p0 = warp; p0 < HS * HS; p0 += NWARP * BATCH) {
        float val[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int pix = p0 + i * NWARP;
            const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
            val[i] = 0.f;
            if (pix < HS * HS && cok && y >= 0 && y < H && x >= 0 && x < W)
                val[i] = __ldg(ub + ((long long)y * W + x) * ld_u + c);
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int pix = p0 + i * NWARP;
            if (pix < HS * HS) h1s[pix][lane] = gelu_f(val[i]);  // gelu(0) = 0 keeps the zero padding
        }
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
        const long long tok0 = ((long long)b * H + y) * W + tx0;
        float gate[TS];
        if (mode == 1) {
#pragma unroll
            for (int lx = 0; lx < TS; ++lx) gate[lx] = (tx0 + lx < W) ? __ldg(u + (tok0 + lx) * ld_u + Ch + c) : 0.f;
        }
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
            if (tx0 + lx < W) {
                float acc = bv;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(win[ky][kx], wgt[ky * 3 + kx], acc);
                float cdf, pdf;
                gelu_parts(acc, cdf, pdf);
                if (v) v[(tok0 + lx) * Ch + c] = v_is_dgelu ? fmaf(acc, pdf, cdf) : acc;
                float o = acc * cdf;
                if (mode == 1) o *= gelu_f(gate[lx]);
                h2[(tok0 + lx) * Ch + c] = rnd ? tf32_round(o) : o;
            }
        }
    }
}

// dv = dh2 * gelu'(v) [* gelu(u2)] ; FRFN additionally du[:, Ch:] = dh2 * gelu(v) * gelu'(u2)
__global__ void __launch_bounds__(256) gelu_gate_bwd_kernel(const float* __restrict__ dh2,
                                                            const float* __restrict__ u, long long ld_u,
                                                            const float* __restrict__ v, float* __restrict__ dv,
                                                            float* __restrict__ du, long long rows, int Ch, int mode) {
    const long long total = rows * Ch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / Ch;
        const int c = (int)(i % Ch);
        const float d = dh2[i], vv = v[i];
        float g = d * gelu_grad_f(vv);
        if (mode == 1) {
            const float u2 = u[r * ld_u + Ch + c];
            g *= gelu_f(u2);
            du[r * ld_u + Ch + c] = d * gelu_f(vv) * gelu_grad_f(u2);
        }
        dv[i] = g;
    }
}

// persistent over tiles: grid = (P, Ch/32); each CTA accumulates dweight/dbias partials in registers
__global__ void __launch_bounds__(DW_THREADS) dwconv_bwd_kernel(const float* __restrict__ dv,
                                                                const float* __restrict__ u, long long ld_u,
                                                                const float* __restrict__ weight,
                                                                float* __restrict__ du, float* __restrict__ partials,
                                                                int B, int H, int W, int Ch, int tiles_x,
                                                                int tiles_per_img, int rnd) {
    __shared__ float dvs[HS * HS][CG];
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
        for (int p0 = warp; p0 < HS * HS; p0 += NWARP * BATCH) {
            float val[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int pix = p0 + i * NWARP;
                const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
                val[i] = 0.f;
                if (pix < HS * HS && cok && y >= 0 && y < H && x >= 0 && x < W)
                    val[i] = __ldg(dv + (base + (long long)y * W + x) * Ch + c);
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int pix = p0 + i * NWARP;
                if (pix < HS * HS) dvs[pix][lane] = val[i];
            }
        }
        __syncthreads();
        if (cok) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int ly = warp * 2 + rr;
                const int y = ty0 + ly;
                if (y >= H) break;
                const long long tok0 = base + (long long)y * W + tx0;
                float uc[TS];
#pragma unroll
                for (int lx = 0; lx < TS; ++lx) uc[lx] = (tx0 + lx < W) ? __ldg(u + (tok0 + lx) * ld_u + c) : 0.f;
                // window of dv around the output pixel: win[a][b] = dv[y-1+a][x-1+b]
                float win[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    win[a][1] = dvs[(ly + a) * HS + 0][lane];
                    win[a][2] = dvs[(ly + a) * HS + 1][lane];
                }
#pragma unroll
                for (int lx = 0; lx < TS; ++lx) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        win[a][0] = win[a][1];
                        win[a][1] = win[a][2];
                        win[a][2] = dvs[(ly + a) * HS + lx + 2][lane];
                    }
                    if (tx0 + lx < W) {
                        const float x = uc[lx];
                        float cdf, pdf;
                        gelu_parts(x, cdf, pdf);
                        const float h1 = x * cdf;
                        float dh1 = 0.f;
                        // v[q] = sum_k h1[q + k - 1] w[k]  =>  h1[p] meets dv[p + 1 - k] with weight w[k]
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float d = win[2 - ky][2 - kx];
                                dh1 = fmaf(d, wgt[ky * 3 + kx], dh1);
                                dwt[ky * 3 + kx] = fmaf(h1, d, dwt[ky * 3 + kx]);
                            }
                        dbs += win[1][1];
                        const float dval = dh1 * (cdf + x * pdf);
                        du[(tok0 + lx) * ld_u + c] = rnd ? tf32_round(dval) : dval;
                    }
                }
            }
        }
    }
    // reduce the 10 per-channel partial sums across the 8 warps (the tile buffer is reused)
    __syncthreads();
    float(*red)[10][CG] = reinterpret_cast<float(*)[10][CG]>(&dvs[0][0]);
#pragma unroll
    for (int k = 0; k < 9; ++k) red[warp][k][lane] = dwt[k];
    red[warp][9][lane] = dbs;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 10 * CG; idx += DW_THREADS) {
        const int k = idx / CG, l = idx % CG;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += red[w][k][l];
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
    int p = uwr_cdiv(4 * uwr_sm_count(), groups);
    if (p > tiles) p = tiles;
    return p < 1 ? 1 : p;
}

}  // namespace

extern "C" int uwr_dwconv_gelu_fwd(const float* u, long long ld_u, const float* weight, const float* bias, float* v,
                                   float* h2, int B, int H, int W, int Ch, int mode, int v_is_dgelu,
                                   uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(u && weight && bias && h2, "uwr_dwconv_gelu_fwd: null pointer");
    UWR_REQUIRE(mode == 0 || mode == 1, "uwr_dwconv_gelu_fwd: mode must be 0 (LeFF) or 1 (FRFN gate)");
    UWR_REQUIRE(ld_u >= (mode == 1 ? 2 * Ch : Ch), "uwr_dwconv_gelu_fwd: ld_u too small");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_dwconv_gelu_fwd: bad batch %d", B);
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    dim3 grid(tx * ty, uwr_cdiv(Ch, CG), B);
    dwconv_fwd_kernel<<<grid, DW_THREADS, 0, stream>>>(u, ld_u, weight, bias, v, h2, H, W, Ch, mode, tx, uwr_round_outputs(),
                                                      v_is_dgelu);
    UWR_CHECK_LAUNCH("dwconv_fwd_kernel");
    return 0;
}

extern "C" int uwr_gelu_gate_bwd(const float* dh2, const float* u, long long ld_u, const float* v, float* dv,
                                 float* du, long long rows, int Ch, int mode, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dh2 && v && dv && (mode == 0 || (u && du)), "uwr_gelu_gate_bwd: null pointer");
    const long long total = rows * Ch;
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    gelu_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dh2, u, ld_u, v, dv, du, rows, Ch, mode);
    UWR_CHECK_LAUNCH("gelu_gate_bwd_kernel");
    return 0;
}

extern "C" size_t uwr_dwconv_gelu_bwd_workspace_bytes(int B, int H, int W, int Ch) {
    return (size_t)bwd_ctas(B, H, W, Ch) * 10 * (size_t)Ch * sizeof(float);
}

extern "C" int uwr_dwconv_gelu_bwd(const float* dv, const float* u, long long ld_u, const float* weight, float* du,
                                   float* dweight, float* dbias, float* workspace, int B, int H, int W, int Ch,
                                   uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dv && u && weight && du && dweight && dbias && workspace, "uwr_dwconv_gelu_bwd: null pointer");
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    const int P = bwd_ctas(B, H, W, Ch);
    dim3 grid(P, uwr_cdiv(Ch, CG));
    dwconv_bwd_kernel<<<grid, DW_THREADS, 0, stream>>>(dv, u, ld_u, weight, du, workspace, B, H, W, Ch, tx, tx * ty,
                                                      uwr_round_outputs());
    UWR_CHECK_LAUNCH("dwconv_bwd_kernel");
    dwconv_param_reduce_kernel<<<uwr_cdiv(10 * Ch, 128), 128, 0, stream>>>(workspace, dweight, dbias, P, Ch);
    UWR_CHECK_LAUNCH("dwconv_param_reduce_kernel");
    return 0;
}
