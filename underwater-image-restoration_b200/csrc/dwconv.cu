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

// thread -> (tile row r, x half xh, 4-channel group c4): 8 outputs x 4 channels, 128-bit accesses only
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
    a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}

__global__ void __launch_bounds__(DW_THREADS) dwconv_fwd_kernel(const float* __restrict__ u, long long ld_u,
                                                                const float* __restrict__ weight,
                                                                const float* __restrict__ bias,
                                                                float* __restrict__ v, float* __restrict__ h2,
                                                                int H, int W, int Ch, int mode, int tiles_x, int rnd,
                                                                int v_is_dgelu) {
    __shared__ __align__(16) float h1s[HS * HS][CG];
    const int tid = threadIdx.x;
    const int c4 = (tid & 7) * 4;
    const int c = blockIdx.y * CG + c4;
    const bool cok = c < Ch;  // Ch is a multiple of 4
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
    const float* ub = u + (long long)b * H * W * ld_u;

    // ---- stage gelu(u) with halo: 324 pixels x 8 float4, all loads of a batch in flight together
    constexpr int NV = (HS * HS * (CG / 4) + DW_THREADS - 1) / DW_THREADS;  // 11
    float4 val[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = tid + i * DW_THREADS;
        const int pix = idx >> 3;
        const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
        val[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pix < HS * HS && cok && y >= 0 && y < H && x >= 0 && x < W)
            val[i] = ld4(ub + ((long long)y * W + x) * ld_u + c);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int idx = tid + i * DW_THREADS;
        const int pix = idx >> 3;
        if (pix < HS * HS)  // gelu(0) = 0 keeps the zero padding; mode 2 = plain depthwise conv
            *reinterpret_cast<float4*>(&h1s[pix][c4]) =
                mode == 2 ? val[i]
                          : make_float4(gelu_f(val[i].x), gelu_f(val[i].y), gelu_f(val[i].z), gelu_f(val[i].w));
    }
    float4 wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        wgt[k] = cok ? make_float4(weight[c * 9 + k], weight[(c + 1) * 9 + k], weight[(c + 2) * 9 + k],
                                   weight[(c + 3) * 9 + k])
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bv = (cok && bias != nullptr) ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const int ly = tid >> 4;            // tile row 0..15
    const int lx0 = ((tid >> 3) & 1) * 8;  // first of 8 columns
    const int y = ty0 + ly;
    if (!cok || y >= H) return;
    const long long tok0 = ((long long)b * H + y) * W + tx0;
    float4 win[3][3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        win[ky][1] = ld4(&h1s[(ly + ky) * HS + lx0][c4]);
        win[ky][2] = ld4(&h1s[(ly + ky) * HS + lx0 + 1][c4]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int lx = lx0 + i;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            win[ky][0] = win[ky][1];
            win[ky][1] = win[ky][2];
            win[ky][2] = ld4(&h1s[(ly + ky) * HS + lx + 2][c4]);
        }
        if (tx0 + lx < W) {
            float4 acc = bv;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) fma4(acc, win[ky][kx], wgt[ky * 3 + kx]);
            float a[4] = {acc.x, acc.y, acc.z, acc.w}, o[4], sv[4];
            float4 gate = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mode == 1) gate = ld4(u + (tok0 + lx) * ld_u + Ch + c);
            const float gt[4] = {gate.x, gate.y, gate.z, gate.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float cdf, pdf;
                gelu_parts(a[e], cdf, pdf);
                sv[e] = v_is_dgelu ? fmaf(a[e], pdf, cdf) : a[e];
                o[e] = mode == 2 ? a[e] : a[e] * cdf;
                if (mode == 1) o[e] *= gelu_f(gt[e]);
                if (rnd) o[e] = tf32_round(o[e]);
            }
            if (v) *reinterpret_cast<float4*>(v + (tok0 + lx) * Ch + c) = make_float4(sv[0], sv[1], sv[2], sv[3]);
            *reinterpret_cast<float4*>(h2 + (tok0 + lx) * Ch + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// dv = dh2 * gelu'(v) [* gelu(u2)] ; FRFN additionally du[:, Ch:] = dh2 * gelu(v) * gelu'(u2)
__global__ void __launch_bounds__(256) gelu_gate_bwd_kernel(const float* __restrict__ dh2,
                                                            const float* __restrict__ u, long long ld_u,
                                                            const float* __restrict__ v, float* __restrict__ dv,
                                                            float* __restrict__ du, long long rows, int Ch, int mode,
                                                            int rnd) {
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
            const float g2 = d * gelu_f(vv) * gelu_grad_f(u2);
            du[r * ld_u + Ch + c] = rnd ? tf32_round(g2) : g2;
        }
        dv[i] = g;
    }
}

// persistent over tiles: grid = (P, Ch/32).  thread -> (tile row, 2-channel group): 16 outputs x 2
// channels with 64-bit accesses (the 4-channel variant needs > 128 registers for the nine dweight
// accumulators and would leave one CTA per SM).  dweight/dbias partials accumulate in registers and
// are reduced across the CTA once, at the end.
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void fma2(float2& a, const float2& x, const float2& w) {
    a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y);
}

__global__ void __launch_bounds__(DW_THREADS, 2) dwconv_bwd_kernel(const float* __restrict__ dv,
                                                                   const float* __restrict__ u, long long ld_u,
                                                                   const float* __restrict__ weight,
                                                                   float* __restrict__ du,
                                                                   float* __restrict__ partials, int B, int H, int W,
                                                                   int Ch, int tiles_x, int tiles_per_img, int rnd,
                                                                   int plain) {
    __shared__ __align__(16) float dvs[HS * HS][CG];
    const int tid = threadIdx.x;
    const int c2 = (tid & 15) * 2;
    const int c = blockIdx.y * CG + c2;
    const bool cok = c < Ch;  // Ch is a multiple of 4
    const int ly = tid >> 4;
    const int s4 = (tid & 7) * 4;  // staging uses 128-bit pieces

    float2 wgt[9], dwt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        wgt[k] = cok ? make_float2(weight[c * 9 + k], weight[(c + 1) * 9 + k]) : make_float2(0.f, 0.f);
        dwt[k] = make_float2(0.f, 0.f);
    }
    float2 dbs = make_float2(0.f, 0.f);
    float2 dus = make_float2(0.f, 0.f);  // column sums of du = bias gradient of the Linear that produced u
    const bool sok = blockIdx.y * CG + s4 < Ch;

    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * TS, tx0 = (tl % tiles_x) * TS;
        const long long base = (long long)b * H * W;
        __syncthreads();
        constexpr int NV = (HS * HS * (CG / 4) + DW_THREADS - 1) / DW_THREADS;  // 11 float4 per thread
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two batches keep the register footprint down
            float4 val[(NV + 1) / 2];
#pragma unroll
            for (int i = 0; i < (NV + 1) / 2; ++i) {
                const int idx = tid + (h * ((NV + 1) / 2) + i) * DW_THREADS;
                const int pix = idx >> 3;
                const int y = ty0 + pix / HS - 1, x = tx0 + pix % HS - 1;
                val[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pix < HS * HS && sok && y >= 0 && y < H && x >= 0 && x < W)
                    val[i] = ld4(dv + (base + (long long)y * W + x) * Ch + blockIdx.y * CG + s4);
            }
#pragma unroll
            for (int i = 0; i < (NV + 1) / 2; ++i) {
                const int idx = tid + (h * ((NV + 1) / 2) + i) * DW_THREADS;
                const int pix = idx >> 3;
                if (pix < HS * HS) *reinterpret_cast<float4*>(&dvs[pix][s4]) = val[i];
            }
        }
        const int y = ty0 + ly;
        const bool rok = cok && y < H;
        const long long tok0 = base + (long long)y * W + tx0;
        __syncthreads();
        if (rok) {
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) {  // two strips of 8 columns
                const int lx0 = hx * 8;
                float2 uc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    uc[i] = (tx0 + lx0 + i < W) ? ld2(u + (tok0 + lx0 + i) * ld_u + c) : make_float2(0.f, 0.f);
                // window of dv around the output pixel: win[a][b] = dv[y-1+a][x-1+b]
                float2 win[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    win[a][1] = ld2(&dvs[(ly + a) * HS + lx0][c2]);
                    win[a][2] = ld2(&dvs[(ly + a) * HS + lx0 + 1][c2]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int lx = lx0 + i;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        win[a][0] = win[a][1];
                        win[a][1] = win[a][2];
                        win[a][2] = ld2(&dvs[(ly + a) * HS + lx + 2][c2]);
                    }
                    if (tx0 + lx < W) {
                        float cdf0 = 1.f, pdf0 = 0.f, cdf1 = 1.f, pdf1 = 0.f;  // plain conv: h1 = u, gelu' = 1
                        if (!plain) {
                            gelu_parts(uc[i].x, cdf0, pdf0);
                            gelu_parts(uc[i].y, cdf1, pdf1);
                        }
                        const float2 h1 = make_float2(uc[i].x * cdf0, uc[i].y * cdf1);
                        float2 dh1 = make_float2(0.f, 0.f);
                        // v[q] = sum_k h1[q + k - 1] w[k]  =>  h1[p] meets dv[p + 1 - k] with weight w[k]
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float2 d = win[2 - ky][2 - kx];
                                fma2(dh1, d, wgt[ky * 3 + kx]);
                                fma2(dwt[ky * 3 + kx], h1, d);
                            }
                        dbs.x += win[1][1].x;
                        dbs.y += win[1][1].y;
                        float2 o = make_float2(dh1.x * fmaf(uc[i].x, pdf0, cdf0), dh1.y * fmaf(uc[i].y, pdf1, cdf1));
                        if (rnd) o = make_float2(tf32_round(o.x), tf32_round(o.y));
                        dus.x += o.x;
                        dus.y += o.y;
                        *reinterpret_cast<float2*>(du + (tok0 + lx) * ld_u + c) = o;
                    }
                }
            }
        }
    }
    // cross-thread reduction: 20 partial sums per thread (10 taps x 2 channels), 16 threads per channel
    // pair; the tile buffer is reused as red[slot][tid]
    __syncthreads();
    float* red = &dvs[0][0];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        red[(k * 2 + 0) * DW_THREADS + tid] = dwt[k].x;
        red[(k * 2 + 1) * DW_THREADS + tid] = dwt[k].y;
    }
    red[18 * DW_THREADS + tid] = dbs.x;
    red[19 * DW_THREADS + tid] = dbs.y;
    red[20 * DW_THREADS + tid] = dus.x;
    red[21 * DW_THREADS + tid] = dus.y;
    __syncthreads();
    for (int idx = tid; idx < 11 * CG; idx += DW_THREADS) {
        const int k = idx / CG, l = idx % CG;  // tap (9 = conv bias, 10 = column sum of du), local channel
        const int grp = l >> 1, e = l & 1;
        float sum = 0.f;
#pragma unroll 8
        for (int j = 0; j < DW_THREADS / 16; ++j) sum += red[(k * 2 + e) * DW_THREADS + j * 16 + grp];
        const int cc = blockIdx.y * CG + l;
        if (cc < Ch) partials[((long long)blockIdx.x * 11 + k) * Ch + cc] = sum;
    }
}

__global__ void dwconv_param_reduce_kernel(const float* __restrict__ partials, float* __restrict__ dweight,
                                           float* __restrict__ dbias, float* __restrict__ du_colsum, int P, int Ch) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 11 * Ch) return;
    const int k = idx / Ch, c = idx % Ch;
    if (k == 10 && du_colsum == nullptr) return;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += partials[((long long)p * 11 + k) * Ch + c];
    if (k < 9) dweight[c * 9 + k] = s;
    else if (k == 9) dbias[c] = s;
    else du_colsum[c] = s;
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
    UWR_REQUIRE(u && weight && h2 && (bias || mode == 2), "uwr_dwconv_gelu_fwd: null pointer");
    UWR_REQUIRE(mode >= 0 && mode <= 2, "uwr_dwconv_gelu_fwd: mode must be 0 (LeFF), 1 (FRFN gate) or 2 (plain)");
    UWR_REQUIRE(ld_u >= (mode == 1 ? 2 * Ch : Ch), "uwr_dwconv_gelu_fwd: ld_u too small");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_dwconv_gelu_fwd: bad batch %d", B);
    UWR_REQUIRE(Ch % 4 == 0 && ld_u % 4 == 0, "uwr_dwconv_gelu_fwd: Ch and ld_u must be multiples of 4");
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
    gelu_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dh2, u, ld_u, v, dv, du, rows, Ch, mode, uwr_round_outputs());
    UWR_CHECK_LAUNCH("gelu_gate_bwd_kernel");
    return 0;
}

extern "C" size_t uwr_dwconv_gelu_bwd_workspace_bytes(int B, int H, int W, int Ch) {
    return (size_t)bwd_ctas(B, H, W, Ch) * 11 * (size_t)Ch * sizeof(float);
}

extern "C" int uwr_dwconv_gelu_bwd(const float* dv, const float* u, long long ld_u, const float* weight, float* du,
                                   float* dweight, float* dbias, float* du_colsum, float* workspace, int B, int H,
                                   int W, int Ch, int plain, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dv && u && weight && du && dweight && dbias && workspace, "uwr_dwconv_gelu_bwd: null pointer");
    UWR_REQUIRE(Ch % 4 == 0 && ld_u % 4 == 0, "uwr_dwconv_gelu_bwd: Ch and ld_u must be multiples of 4");
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    const int P = bwd_ctas(B, H, W, Ch);
    dim3 grid(P, uwr_cdiv(Ch, CG));
    dwconv_bwd_kernel<<<grid, DW_THREADS, 0, stream>>>(dv, u, ld_u, weight, du, workspace, B, H, W, Ch, tx, tx * ty,
                                                      uwr_round_outputs(), plain);
    UWR_CHECK_LAUNCH("dwconv_bwd_kernel");
    dwconv_param_reduce_kernel<<<uwr_cdiv(11 * Ch, 128), 128, 0, stream>>>(workspace, dweight, dbias, du_colsum, P, Ch);
    UWR_CHECK_LAUNCH("dwconv_param_reduce_kernel");
    return 0;
}
