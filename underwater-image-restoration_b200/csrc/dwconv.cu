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
#include <cuda_fp16.h>

#include "uwr_common.cuh"
#include "uwr_tma.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int TS = 16;          // tile side
constexpr int HS = TS + 2;      // halo side
constexpr int CG = 32;          // channels per CTA
constexpr int DW_THREADS = 256; // 8 warps, 2 tile rows each

// thread -> (tile row r, x half xh, 4-channel group c4): 8 outputs x 4 channels, 128-bit accesses only
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
    // scalar on purpose: the forward kernels measured SLOWER with packed FFMA2 / FMUL2 (0.560 vs 0.523 ms at
    // B16 H256 Ch256, same box) while the backward gains 5 % (0.590 vs 0.619 ms) -- see fma2 below
    a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}

// Halo tile (18 x 18 pixels x 32 channels, 41 472 B) of tile `t` -> shared memory by ONE TMA
// instruction: 4-D tensor map over (C, W, H, B), box (32, 18, 18, 1) starting at (cbase, tx0-1, ty0-1, b);
// the hardware zero-fills everything outside the image / beyond Ch, which is exactly the conv padding.
constexpr int TILE_BYTES = HS * HS * CG * (int)sizeof(float);
static_assert(TILE_BYTES % 128 == 0, "tile buffers must stay 128-byte aligned");

__device__ __forceinline__ void dw_tma_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int tile,
                                            int tiles_per_img, int tiles_x, int cbase, int bytes = TILE_BYTES) {
    const int b = tile / tiles_per_img, tl = tile - b * tiles_per_img;
    const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
    uwr_tma::mbar_expect_tx(bar, bytes);
    uwr_tma::tma_load_4d(dst, map, bar, cbase, tx * TS - 1, ty * TS - 1, b);
}

// persistent: grid = (P, Ch/32).  The halo tile of the NEXT (image, tile) is in flight (TMA, mbarrier)
// while the current one is convolved.  MODE 0 LeFF, 1 FRFN gate, 2 plain conv.  FAST = training hot
// path (v stored as gelu'(v), h2 rounded to TF32, full 16x16x32 tiles): no flags or bounds in the loop.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_fwd_kernel(const __grid_constant__ CUtensorMap map_u, const float* __restrict__ u, long long ld_u,
                  const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ v,
                  float* __restrict__ h2, int B, int H, int W, int Ch, int tiles_x, int tiles_per_img, int rnd,
                  int v_is_dgelu) {
    uwr_pdl_enter();
    extern __shared__ __align__(128) unsigned char dw_raw[];
    // 128-byte alignment for the TMA destination, computed on the shared-window offset so that the
    // compiler keeps LDS/STS (a generic uintptr_t round-up turned every access into LD.E/ST.E)
    float* smem = reinterpret_cast<float*>(dw_raw + ((128u - (uwr_tma::smem_u32(dw_raw) & 127u)) & 127u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (TILE_BYTES / 4));
    const int tid = threadIdx.x;
    const int c4 = (tid & 7) * 4;
    const int cbase = blockIdx.y * CG;
    const int c = cbase + c4;
    const bool cok = FAST || c < Ch;  // Ch is a multiple of 4
    const int total = B * tiles_per_img;
    int tile = blockIdx.x;

    if (tid == 0) {
        uwr_tma::mbar_init(&bars[0], 1);
        uwr_tma::mbar_init(&bars[1], 1);
        uwr_tma::mbar_init_fence();
        uwr_tma::tma_prefetch_map(&map_u);
    }
    __syncthreads();
    if (tid == 0 && tile < total) dw_tma_tile(smem, &map_u, &bars[0], tile, tiles_per_img, tiles_x, cbase);

    float4 wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        wgt[k] = cok ? make_float4(weight[c * 9 + k], weight[(c + 1) * 9 + k], weight[(c + 2) * 9 + k],
                                   weight[(c + 3) * 9 + k])
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bv = (cok && bias != nullptr) ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int ly = tid >> 4;               // tile row 0..15
    const int lx0 = ((tid >> 3) & 1) * 8;  // first of 8 columns
    const bool vdg = FAST || v_is_dgelu, rn = FAST || rnd, hasv = FAST || v != nullptr;

    for (int it = 0; tile < total; tile += gridDim.x, ++it) {
        float* cur = smem + (it & 1) * (TILE_BYTES / 4);
        const int ntile = tile + gridDim.x;
        // the other buffer was released by the __syncthreads that ended the previous iteration
        if (tid == 0 && ntile < total)
            dw_tma_tile(smem + ((it + 1) & 1) * (TILE_BYTES / 4), &map_u, &bars[(it + 1) & 1], ntile, tiles_per_img,
                        tiles_x, cbase);
        uwr_tma::mbar_wait(&bars[it & 1], (it >> 1) & 1);
        if (MODE != 2) {  // GELU in place, once per staged element (gelu(0) = 0 keeps the zero padding)
            float4* p4 = reinterpret_cast<float4*>(cur) + tid;
#pragma unroll
            for (int i = 0; i < 10; ++i) {  // 18*18*8 = 2592 float4 = 10 * 256 + 32
                float4 t4 = p4[i * DW_THREADS];
                p4[i * DW_THREADS] = make_float4(gelu_f(t4.x), gelu_f(t4.y), gelu_f(t4.z), gelu_f(t4.w));
            }
            if (tid < 32) {
                float4 t4 = p4[10 * DW_THREADS];
                p4[10 * DW_THREADS] = make_float4(gelu_f(t4.x), gelu_f(t4.y), gelu_f(t4.z), gelu_f(t4.w));
            }
            uwr_tma::fence_proxy_async_smem();  // these generic writes precede the TMA that reuses the buffer
            __syncthreads();
        }
        const int b = tile / tiles_per_img, tl = tile - b * tiles_per_img;
        const int ty0 = (tl / tiles_x) * TS, tx0 = (tl % tiles_x) * TS;
        const int y = ty0 + ly;
        if (FAST || (cok && y < H)) {
            const long long tok0 = ((long long)b * H + y) * W + tx0 + lx0;
            float* h2p = h2 + tok0 * Ch + c;
            float* vp = hasv ? v + tok0 * Ch + c : nullptr;
            const float* gp = u + tok0 * ld_u + Ch + c;  // FRFN gate half (MODE 1)
            const float* sp = cur + (ly * HS + lx0) * CG + c4;
            float4 col[3][3];  // col[ky][j % 3] = staged pixel (ly + ky, lx0 + j): rotating registers, no moves
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                col[ky][0] = ld4(sp + (ky * HS + 0) * CG);
                col[ky][1] = ld4(sp + (ky * HS + 1) * CG);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) col[ky][(i + 2) % 3] = ld4(sp + (ky * HS + i + 2) * CG);
                if (FAST || tx0 + lx0 + i < W) {
                    float4 acc = bv;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) fma4(acc, col[ky][(i + kx) % 3], wgt[ky * 3 + kx]);
                    float a[4] = {acc.x, acc.y, acc.z, acc.w}, o[4], sv[4];
                    float gt[4] = {0.f, 0.f, 0.f, 0.f};
                    if (MODE == 1) {
                        const float4 gate = ld4(gp + (long long)i * ld_u);
                        gt[0] = gate.x; gt[1] = gate.y; gt[2] = gate.z; gt[3] = gate.w;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (MODE == 2) {
                            sv[e] = a[e];
                            o[e] = a[e];
                        } else {
                            float cdf, pdf;
                            gelu_parts(a[e], cdf, pdf);
                            sv[e] = vdg ? fmaf(a[e], pdf, cdf) : a[e];
                            o[e] = a[e] * cdf;
                            if (MODE == 1) o[e] *= gelu_f(gt[e]);
                        }
                        if (rn) o[e] = tf32_round(o[e]);
                    }
                    if (hasv) *reinterpret_cast<float4*>(vp + (long long)i * Ch) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                    *reinterpret_cast<float4*>(h2p + (long long)i * Ch) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        __syncthreads();  // everyone is done with `cur` before the TMA of iteration it+1 overwrites it
    }
}

// LeFF training hot path with fp16 storage of the 4C-wide tensors (DESIGN.md §3 "half storage"): u arrives as
// __half (written by the linear1 GEMM epilogue), gelu'(v) leaves as __half (read back by the linear2 data-gradient
// epilogue); h2 stays fp32 (TF32-rounded A operand of linear2).  Same tiling as dwconv_fwd_kernel<0, true>:
// the halo tile of halves lands by ONE TMA instruction (20 736 B, half the bytes) and is expanded to fp32 in
// shared memory by the GELU pass that the fp32 kernel runs in place.  Full tiles only (H, W % 16 == 0, Ch % 32 == 0).
constexpr int TILE_BYTES_H = HS * HS * CG * 2;
static_assert(TILE_BYTES_H % 128 == 0, "half tile buffers must stay 128-byte aligned");

__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_fwd_half_kernel(const __grid_constant__ CUtensorMap map_u, const float* __restrict__ weight,
                       const float* __restrict__ bias, __half* __restrict__ v, float* __restrict__ h2, int B, int H, int W,
                       int Ch, int tiles_x, int tiles_per_img) {
    uwr_pdl_enter();
    extern __shared__ __align__(128) unsigned char dw_raw[];
    unsigned char* base = dw_raw + ((128u - (uwr_tma::smem_u32(dw_raw) & 127u)) & 127u);
    float* work = reinterpret_cast<float*>(base);                                  // gelu(u) tile, fp32
    unsigned char* hbuf = base + TILE_BYTES;                                       // two __half halo tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + 2 * TILE_BYTES_H);
    const int tid = threadIdx.x;
    const int c4 = (tid & 7) * 4;
    const int cbase = blockIdx.y * CG;
    const int c = cbase + c4;
    const int total = B * tiles_per_img;
    int tile = blockIdx.x;

    if (tid == 0) {
        uwr_tma::mbar_init(&bars[0], 1);
        uwr_tma::mbar_init(&bars[1], 1);
        uwr_tma::mbar_init_fence();
        uwr_tma::tma_prefetch_map(&map_u);
    }
    __syncthreads();
    if (tid == 0 && tile < total) dw_tma_tile(hbuf, &map_u, &bars[0], tile, tiles_per_img, tiles_x, cbase, TILE_BYTES_H);

    float4 wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        wgt[k] = make_float4(weight[c * 9 + k], weight[(c + 1) * 9 + k], weight[(c + 2) * 9 + k], weight[(c + 3) * 9 + k]);
    const float4 bv = ld4(bias + c);
    const int ly = tid >> 4;
    const int lx0 = ((tid >> 3) & 1) * 8;

    for (int it = 0; tile < total; tile += gridDim.x, ++it) {
        const uint2* cur = reinterpret_cast<const uint2*>(hbuf + (it & 1) * TILE_BYTES_H);
        const int ntile = tile + gridDim.x;
        // the other half buffer was drained by the GELU pass of the previous iteration (and a __syncthreads since)
        if (tid == 0 && ntile < total)
            dw_tma_tile(hbuf + ((it + 1) & 1) * TILE_BYTES_H, &map_u, &bars[(it + 1) & 1], ntile, tiles_per_img, tiles_x,
                        cbase, TILE_BYTES_H);
        uwr_tma::mbar_wait(&bars[it & 1], (it >> 1) & 1);
        {   // fp16 -> fp32 + GELU, once per staged element (gelu(0) = 0 keeps the zero padding)
            float4* w4 = reinterpret_cast<float4*>(work) + tid;
            const uint2* s2 = cur + tid;
#pragma unroll
            for (int i = 0; i < 10; ++i) {  // 18*18*8 = 2592 groups of 4 channels = 10 * 256 + 32
                const uint2 raw = s2[i * DW_THREADS];
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                w4[i * DW_THREADS] = make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(b2.x), gelu_f(b2.y));
            }
            if (tid < 32) {
                const uint2 raw = s2[10 * DW_THREADS];
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                w4[10 * DW_THREADS] = make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(b2.x), gelu_f(b2.y));
            }
            __syncthreads();
        }
        const int b = tile / tiles_per_img, tl = tile - b * tiles_per_img;
        const int ty0 = (tl / tiles_x) * TS, tx0 = (tl % tiles_x) * TS;
        const long long tok0 = ((long long)b * H + ty0 + ly) * W + tx0 + lx0;
        float* h2p = h2 + tok0 * Ch + c;
        __half* vp = v + tok0 * Ch + c;
        const float* sp = work + (ly * HS + lx0) * CG + c4;
        float4 col[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            col[ky][0] = ld4(sp + (ky * HS + 0) * CG);
            col[ky][1] = ld4(sp + (ky * HS + 1) * CG);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) col[ky][(i + 2) % 3] = ld4(sp + (ky * HS + i + 2) * CG);
            float4 acc = bv;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) fma4(acc, col[ky][(i + kx) % 3], wgt[ky * 3 + kx]);
            const float a[4] = {acc.x, acc.y, acc.z, acc.w};
            float o[4], sv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float cdf, pdf;
                gelu_parts(a[e], cdf, pdf);
                sv[e] = fmaf(a[e], pdf, cdf);          // gelu'(v) in [-0.13, 1.13]: fp16 keeps it to 2^-11
                o[e] = tf32_round(a[e] * cdf);
            }
            const __half2 s01 = __floats2half2_rn(sv[0], sv[1]), s23 = __floats2half2_rn(sv[2], sv[3]);
            uint2 raw;
            raw.x = *reinterpret_cast<const uint32_t*>(&s01);
            raw.y = *reinterpret_cast<const uint32_t*>(&s23);
            *reinterpret_cast<uint2*>(vp + (long long)i * Ch) = raw;
            *reinterpret_cast<float4*>(h2p + (long long)i * Ch) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();  // everyone is done with `work` before the next iteration's GELU pass rewrites it
    }
}

// dv = dh2 * gelu'(v) [* gelu(u2)] ; FRFN additionally du[:, Ch:] = dh2 * gelu(v) * gelu'(u2)
__global__ void __launch_bounds__(256) gelu_gate_bwd_kernel(const float* __restrict__ dh2,
                                                            const float* __restrict__ u, long long ld_u,
                                                            const float* __restrict__ v, float* __restrict__ dv,
                                                            float* __restrict__ du, long long rows, int Ch, int mode,
                                                            int rnd) {
    uwr_pdl_enter();
    const int c4n = Ch >> 2;  // Ch is a multiple of 4: 128-bit accesses
    const long long total = rows * c4n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4n;
        const int c = (int)(i - r * c4n) * 4;
        const float4 d4 = ld4(dh2 + r * Ch + c), v4 = ld4(v + r * Ch + c);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
        float g[4], g2[4];
        float u2[4] = {0.f, 0.f, 0.f, 0.f};
        if (mode == 1) {
            const float4 t = ld4(u + r * ld_u + Ch + c);
            u2[0] = t.x; u2[1] = t.y; u2[2] = t.z; u2[3] = t.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float cdf, pdf;
            gelu_parts(vv[e], cdf, pdf);
            g[e] = d[e] * fmaf(vv[e], pdf, cdf);          // dh2 * gelu'(v)
            if (mode == 1) {
                float cdf2, pdf2;
                gelu_parts(u2[e], cdf2, pdf2);
                g[e] *= u2[e] * cdf2;                     // ... * gelu(u2)
                g2[e] = d[e] * (vv[e] * cdf) * fmaf(u2[e], pdf2, cdf2);   // dh2 * gelu(v) * gelu'(u2)
                if (rnd) g2[e] = tf32_round(g2[e]);
            }
        }
        if (mode == 1) *reinterpret_cast<float4*>(du + r * ld_u + Ch + c) = make_float4(g2[0], g2[1], g2[2], g2[3]);
        *reinterpret_cast<float4*>(dv + r * Ch + c) = make_float4(g[0], g[1], g[2], g[3]);
    }
}

// GDFN gate (SpectralTransformer.py:126-129): out = gelu(t[:, :h]) * t[:, h:2h] on token slabs (row stride ld)
__global__ void __launch_bounds__(256) gelu_mul_fwd_kernel(const float* __restrict__ t, long long ld,
                                                           float* __restrict__ out, long long rows, int h4, int rnd) {
    uwr_pdl_enter();
    const long long total = rows * h4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / h4;
        const int c = (int)(i % h4) * 4;
        const float4 a = ld4(t + r * ld + c), b = ld4(t + r * ld + h4 * 4 + c);
        float4 o = make_float4(gelu_f(a.x) * b.x, gelu_f(a.y) * b.y, gelu_f(a.z) * b.z, gelu_f(a.w) * b.w);
        if (rnd) o = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
        reinterpret_cast<float4*>(out)[i] = o;
    }
}
// dt[:, :h] = dout * t2 * gelu'(t1) ; dt[:, h:] = dout * gelu(t1)
__global__ void __launch_bounds__(256) gelu_mul_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ t,
                                                           long long ld, float* __restrict__ dt, long long rows,
                                                           int h4) {
    uwr_pdl_enter();
    const long long total = rows * h4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / h4;
        const int c = (int)(i % h4) * 4;
        const float4 a = ld4(t + r * ld + c), b = ld4(t + r * ld + h4 * 4 + c);
        const float4 d = reinterpret_cast<const float4*>(dout)[i];
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, dv[4] = {d.x, d.y, d.z, d.w};
        float g1[4], g2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float cdf, pdf;
            gelu_parts(av[e], cdf, pdf);
            g1[e] = dv[e] * bv[e] * fmaf(av[e], pdf, cdf);
            g2[e] = dv[e] * av[e] * cdf;
        }
        *reinterpret_cast<float4*>(dt + r * ld + c) = make_float4(g1[0], g1[1], g1[2], g1[3]);
        *reinterpret_cast<float4*>(dt + r * ld + h4 * 4 + c) = make_float4(g2[0], g2[1], g2[2], g2[3]);
    }
}

// persistent over tiles: grid = (P, Ch/32).  thread -> (tile row, 2-channel group): 16 outputs x 2
// channels with 64-bit accesses (the 4-channel variant needs > 128 registers for the nine dweight
// accumulators and would leave one CTA per SM).  dweight/dbias partials accumulate in registers and
// are reduced across the CTA once, at the end.
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void fma2(float2& a, const float2& x, const float2& w) {
    a = ffma2(x, w, a);   // packed FFMA2: the backward kernel's 18 tap products per pixel pair, half the issue slots
}

// UHALF: u is __half (fp16 storage of the linear1 output, see dwconv_fwd_half_kernel); ld_u counts elements of u and
// is also the row stride of the fp32 du.
// raw 8-byte (fp32 pair) or 4-byte (fp16 pair, in .x) load; converted where it is used so that the eight loads of a
// strip are in flight together
template <bool UHALF>
__device__ __forceinline__ float2 ld_u2(const float* p) {
    if (UHALF) return make_float2(*p, 0.f);
    return *reinterpret_cast<const float2*>(p);
}
template <bool UHALF>
__device__ __forceinline__ float2 cvt_u2(float2 raw) {
    if (UHALF) {
        const uint32_t bits = __float_as_uint(raw.x);
        return __half22float2(*reinterpret_cast<const __half2*>(&bits));
    }
    return raw;
}

template <bool PLAIN, bool FAST, bool UHALF = false>
__global__ void __launch_bounds__(DW_THREADS, 2)
dwconv_bwd_kernel(const __grid_constant__ CUtensorMap map_dv, const float* __restrict__ u, long long ld_u,
                  const float* __restrict__ weight, float* __restrict__ du, float* __restrict__ partials, int B, int H,
                  int W, int Ch, int tiles_x, int tiles_per_img, int rnd) {
    uwr_pdl_enter();
    extern __shared__ __align__(128) unsigned char dw_raw[];
    float* smem = reinterpret_cast<float*>(dw_raw + ((128u - (uwr_tma::smem_u32(dw_raw) & 127u)) & 127u));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (TILE_BYTES / 4));
    const int tid = threadIdx.x;
    const int c2 = (tid & 15) * 2;
    const int cbase = blockIdx.y * CG;
    const int c = cbase + c2;
    const bool cok = FAST || c < Ch;  // Ch is a multiple of 4
    const int ly = tid >> 4;
    const bool rn = FAST || rnd;
    const int total = B * tiles_per_img;
    int tile = blockIdx.x;

    if (tid == 0) {
        uwr_tma::mbar_init(&bars[0], 1);
        uwr_tma::mbar_init(&bars[1], 1);
        uwr_tma::mbar_init_fence();
        uwr_tma::tma_prefetch_map(&map_dv);
    }
    __syncthreads();
    if (tid == 0 && tile < total) dw_tma_tile(smem, &map_dv, &bars[0], tile, tiles_per_img, tiles_x, cbase);

    float2 wgt[9], dwt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        wgt[k] = cok ? make_float2(weight[c * 9 + k], weight[(c + 1) * 9 + k]) : make_float2(0.f, 0.f);
        dwt[k] = make_float2(0.f, 0.f);
    }
    float2 dbs = make_float2(0.f, 0.f);
    float2 dus = make_float2(0.f, 0.f);  // column sums of du = bias gradient of the Linear that produced u

    for (int it = 0; tile < total; tile += gridDim.x, ++it) {
        const float* dvs = smem + (it & 1) * (TILE_BYTES / 4);
        const int ntile = tile + gridDim.x;
        if (tid == 0 && ntile < total)
            dw_tma_tile(smem + ((it + 1) & 1) * (TILE_BYTES / 4), &map_dv, &bars[(it + 1) & 1], ntile, tiles_per_img,
                        tiles_x, cbase);
        const int b = tile / tiles_per_img, tl = tile - b * tiles_per_img;
        const int ty0 = (tl / tiles_x) * TS, tx0 = (tl % tiles_x) * TS;
        const int y = ty0 + ly;
        const bool rok = FAST || (cok && y < H);
        const long long tok0 = ((long long)b * H + y) * W + tx0;
        // element offsets; with UHALF the byte address is half of what the float pointer arithmetic would give
        const long long uoff = tok0 * ld_u + c;
        const float* up = UHALF ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(u) + uoff) : u + uoff;
        const long long ustep = UHALF ? ld_u / 2 : ld_u;   // pointer steps in floats per pixel (ld_u is even)
        float* dup = du + uoff;
        // the first strip's u values are requested before waiting for the halo tile
        // both strips' u values are requested before waiting for the halo tile (fp16 u: 16 four-byte loads in flight; the
        // second strip used to issue its loads after the first strip's arithmetic: long_scoreboard was 36 % of the stalls)
        float2 uc[UHALF ? 16 : 8];
#pragma unroll
        for (int i = 0; i < (UHALF ? 16 : 8); ++i)
            uc[i] = (rok && (FAST || tx0 + i < W)) ? ld_u2<UHALF>(up + (long long)i * ustep) : make_float2(0.f, 0.f);
        uwr_tma::mbar_wait(&bars[it & 1], (it >> 1) & 1);
        if (rok) {
            const float* sp = dvs + (ly * HS) * CG + c2;
#pragma unroll
            for (int hx = 0; hx < 2; ++hx) {  // two strips of 8 columns
                const int lx0 = hx * 8;
                if (hx == 1 && !UHALF) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        uc[i] = (FAST || tx0 + 8 + i < W) ? ld_u2<UHALF>(up + (long long)(8 + i) * ustep) : make_float2(0.f, 0.f);
                }
                // col[a][j % 3] = dv at tile pixel (ly - 1 + a, lx0 - 1 + j): rotating registers, no moves
                float2 col[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    col[a][0] = ld2(sp + (a * HS + lx0 + 0) * CG);
                    col[a][1] = ld2(sp + (a * HS + lx0 + 1) * CG);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) col[a][(i + 2) % 3] = ld2(sp + (a * HS + lx0 + i + 2) * CG);
                    if (FAST || tx0 + lx0 + i < W) {
                        float2 cdf = make_float2(1.f, 1.f), pdf = make_float2(0.f, 0.f);  // plain conv: h1 = u, gelu' = 1
                        const float2 uv = cvt_u2<UHALF>(uc[UHALF ? lx0 + i : i]);
                        if (!PLAIN) gelu_parts2(uv, cdf, pdf);
                        const float2 h1 = PLAIN ? uv : fmul2(uv, cdf);
                        float2 dh1 = make_float2(0.f, 0.f);
                        // v[q] = sum_k h1[q + k - 1] w[k]  =>  h1[p] meets dv[p + 1 - k] with weight w[k]
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float2 d = col[2 - ky][(i + 2 - kx) % 3];
                                fma2(dh1, d, wgt[ky * 3 + kx]);
                                fma2(dwt[ky * 3 + kx], h1, d);
                            }
                        const float2 ctr = col[1][(i + 1) % 3];
                        dbs = fadd2(dbs, ctr);
                        float2 o = PLAIN ? dh1 : fmul2(dh1, ffma2(uv, pdf, cdf));
                        if (rn) o = make_float2(tf32_round(o.x), tf32_round(o.y));
                        dus = fadd2(dus, o);
                        *reinterpret_cast<float2*>(dup + (long long)(lx0 + i) * ld_u) = o;
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with this buffer before the next TMA overwrites it
    }
    // cross-thread reduction: 22 partial sums per thread (11 slots x 2 channels), 16 threads per channel
    // pair; the tile buffers (2 x 10 368 floats >= 22 x 256) are reused as red[slot][tid]
    float* red = smem;
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
    uwr_pdl_enter();
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

constexpr int DW_SMEM_BYTES = 2 * TILE_BYTES + 128 + 16;  // two halo buffers + alignment slack + 2 mbarriers

// (C, W, H, B) view of a token tensor with row stride ld; box = one halo tile of 32 channels
int dw_halo_map(CUtensorMap* m, const float* base, long long ld, int B, int H, int W, int Ch) {
    const long long dims[4] = {Ch, W, H, B};
    const long long strides[3] = {ld, (long long)W * ld, (long long)H * W * ld};
    const int box[4] = {CG, HS, HS, 1};
    return uwr_tma::encode_f32(m, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int bwd_ctas(int B, int H, int W, int Ch) {
    const int tiles = B * uwr_cdiv(H, TS) * uwr_cdiv(W, TS);
    const int groups = uwr_cdiv(Ch, CG);
    // 2 CTAs/SM (two 41 KB halo buffers each), ONE resident wave: rounded DOWN.  Rounded up, 16 channel groups got 19 CTAs
    // each = 304 CTAs on 296 slots, and the 8 left-over CTAs walked their 54 tiles after everyone else had finished
    // (Ch = 512 / 1024 / 2048 ran at 2.6 TB/s where Ch = 256, 37 x 8 = 296, reached 3.9)
    int p = (2 * uwr_sm_count()) / groups;
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
    CUtensorMap map_u;
    if (dw_halo_map(&map_u, u, ld_u, B, H, W, Ch) != 0) return -3;
    const bool fast = v != nullptr && v_is_dgelu && uwr_round_outputs() && H % TS == 0 && W % TS == 0 && Ch % CG == 0;
    dim3 grid(bwd_ctas(B, H, W, Ch), uwr_cdiv(Ch, CG));
#define DW_FWD(M, F)                                                                                              \
    do {                                                                                                          \
        static bool configured = false;                                                                           \
        if (!configured) {                                                                                        \
            UWR_CUDA(cudaFuncSetAttribute(dwconv_fwd_kernel<M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          DW_SMEM_BYTES));                                                        \
            configured = true;                                                                                    \
        }                                                                                                         \
        (void)uwr_launch_pdl(dwconv_fwd_kernel<M, F>, dim3(grid), dim3(DW_THREADS), DW_SMEM_BYTES, stream,                                      \
            map_u, u, ld_u, weight, bias, v, h2, B, H, W, Ch, tx, tx * ty, uwr_round_outputs(), v_is_dgelu);      \
    } while (0)
    if (mode == 0) {
        if (fast) DW_FWD(0, true);
        else DW_FWD(0, false);
    } else if (mode == 1) {
        if (fast) DW_FWD(1, true);
        else DW_FWD(1, false);
    } else {
        DW_FWD(2, false);
    }
#undef DW_FWD
    UWR_CHECK_LAUNCH("dwconv_fwd_kernel");
    return 0;
}

// fp16-storage variant of the LeFF forward (mode 0, training): u is __half (B*H*W, Ch) dense, gelu'(v) is written as
// __half, h2 as TF32-rounded fp32.  Served when H, W are multiples of 16 and Ch of 32 (uwr_dwconv_half_supported).
extern "C" int uwr_dwconv_half_supported(int H, int W, int Ch) {
    return H > 0 && W > 0 && Ch > 0 && H % TS == 0 && W % TS == 0 && Ch % CG == 0;
}

extern "C" int uwr_dwconv_gelu_fwd_half(const void* u_half, const float* weight, const float* bias, void* dgelu_half,
                                        float* h2, int B, int H, int W, int Ch, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(u_half && weight && bias && dgelu_half && h2, "uwr_dwconv_gelu_fwd_half: null pointer");
    UWR_REQUIRE(uwr_dwconv_half_supported(H, W, Ch), "uwr_dwconv_gelu_fwd_half: needs H, W %% 16 == 0 and Ch %% 32 == 0");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_dwconv_gelu_fwd_half: bad batch %d", B);
    UWR_REQUIRE(uwr_round_outputs(), "uwr_dwconv_gelu_fwd_half: single-pass (tf32) mode only");
    const int tx = W / TS, ty = H / TS;
    CUtensorMap map_u;
    {
        const long long dims[4] = {Ch, W, H, B};
        const long long strides[3] = {Ch, (long long)W * Ch, (long long)H * W * Ch};
        const int box[4] = {CG, HS, HS, 1};
        if (uwr_tma::encode_f16(&map_u, u_half, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) != 0) return -3;
    }
    constexpr int SMEM = TILE_BYTES + 2 * TILE_BYTES_H + 128 + 16;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(dwconv_fwd_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    dim3 grid(bwd_ctas(B, H, W, Ch), Ch / CG);
    (void)uwr_launch_pdl(dwconv_fwd_half_kernel, dim3(grid), dim3(DW_THREADS), SMEM, stream, map_u, weight, bias, (__half*)dgelu_half, h2, B, H, W, Ch, tx,
                                                              tx * ty);
    UWR_CHECK_LAUNCH("dwconv_fwd_half_kernel");
    return 0;
}

extern "C" int uwr_gelu_gate_bwd(const float* dh2, const float* u, long long ld_u, const float* v, float* dv,
                                 float* du, long long rows, int Ch, int mode, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dh2 && v && dv && (mode == 0 || (u && du)), "uwr_gelu_gate_bwd: null pointer");
    UWR_REQUIRE(Ch % 4 == 0 && (mode == 0 || ld_u % 4 == 0), "uwr_gelu_gate_bwd: Ch and ld_u must be multiples of 4");
    const long long total = rows * (Ch / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    (void)uwr_launch_pdl(gelu_gate_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, dh2, u, ld_u, v, dv, du, rows, Ch, mode, uwr_round_outputs());
    UWR_CHECK_LAUNCH("gelu_gate_bwd_kernel");
    return 0;
}

extern "C" int uwr_gelu_mul_fwd(const float* t, long long ld, float* out, long long rows, int h, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(t && out && h % 4 == 0 && ld % 4 == 0 && ld >= 2 * h, "uwr_gelu_mul_fwd: bad args");
    if (rows == 0) return 0;
    long long blocks = (rows * (h / 4) + 255) / 256;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    (void)uwr_launch_pdl(gelu_mul_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, t, ld, out, rows, h / 4, uwr_round_outputs());
    UWR_CHECK_LAUNCH("gelu_mul_fwd_kernel");
    return 0;
}

extern "C" int uwr_gelu_mul_bwd(const float* dout, const float* t, long long ld, float* dt, long long rows, int h,
                                uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dout && t && dt && h % 4 == 0 && ld % 4 == 0 && ld >= 2 * h, "uwr_gelu_mul_bwd: bad args");
    if (rows == 0) return 0;
    long long blocks = (rows * (h / 4) + 255) / 256;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    (void)uwr_launch_pdl(gelu_mul_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, dout, t, ld, dt, rows, h / 4);
    UWR_CHECK_LAUNCH("gelu_mul_bwd_kernel");
    return 0;
}

extern "C" size_t uwr_dwconv_gelu_bwd_workspace_bytes(int B, int H, int W, int Ch) {
    return (size_t)bwd_ctas(B, H, W, Ch) * 11 * (size_t)Ch * sizeof(float);
}

static int dwconv_bwd_impl(const float* dv, const float* u, long long ld_u, const float* weight, float* du, float* dweight,
                           float* dbias, float* du_colsum, float* workspace, int B, int H, int W, int Ch, int plain,
                           int u_half, cudaStream_t stream);

extern "C" int uwr_dwconv_gelu_bwd(const float* dv, const float* u, long long ld_u, const float* weight, float* du,
                                   float* dweight, float* dbias, float* du_colsum, float* workspace, int B, int H,
                                   int W, int Ch, int plain, uwr_stream_t stream_) {
    return dwconv_bwd_impl(dv, u, ld_u, weight, du, dweight, dbias, du_colsum, workspace, B, H, W, Ch, plain, 0,
                           (cudaStream_t)stream_);
}

// same with u stored as __half (B*H*W, Ch) dense (fp16 storage of the linear1 output); du is fp32 (B*H*W, Ch)
extern "C" int uwr_dwconv_gelu_bwd_half(const float* dv, const void* u_half, const float* weight, float* du,
                                        float* dweight, float* dbias, float* du_colsum, float* workspace, int B, int H,
                                        int W, int Ch, uwr_stream_t stream_) {
    UWR_REQUIRE(uwr_dwconv_half_supported(H, W, Ch) && uwr_round_outputs(),
                "uwr_dwconv_gelu_bwd_half: needs H, W %% 16 == 0, Ch %% 32 == 0 and single-pass (tf32) mode");
    return dwconv_bwd_impl(dv, (const float*)u_half, Ch, weight, du, dweight, dbias, du_colsum, workspace, B, H, W, Ch, 0, 1,
                           (cudaStream_t)stream_);
}

static int dwconv_bwd_impl(const float* dv, const float* u, long long ld_u, const float* weight, float* du, float* dweight,
                           float* dbias, float* du_colsum, float* workspace, int B, int H, int W, int Ch, int plain,
                           int u_half, cudaStream_t stream) {
    UWR_REQUIRE(dv && u && weight && du && dweight && dbias && workspace, "uwr_dwconv_gelu_bwd: null pointer");
    UWR_REQUIRE(Ch % 4 == 0 && ld_u % 4 == 0, "uwr_dwconv_gelu_bwd: Ch and ld_u must be multiples of 4");
    const int tx = uwr_cdiv(W, TS), ty = uwr_cdiv(H, TS);
    const int P = bwd_ctas(B, H, W, Ch);
    dim3 grid(P, uwr_cdiv(Ch, CG));
    CUtensorMap map_dv;
    if (dw_halo_map(&map_dv, dv, Ch, B, H, W, Ch) != 0) return -3;
    const bool fast = uwr_round_outputs() && H % TS == 0 && W % TS == 0 && Ch % CG == 0;
#define DW_BWD(PL, F, ...)                                                                                        \
    do {                                                                                                          \
        static bool configured = false;                                                                           \
        auto kern = dwconv_bwd_kernel<PL, F, ##__VA_ARGS__>;                                                      \
        if (!configured) {                                                                                        \
            UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));     \
            configured = true;                                                                                    \
        }                                                                                                         \
        (void)uwr_launch_pdl(kern, dim3(grid), dim3(DW_THREADS), DW_SMEM_BYTES, stream, map_dv, u, ld_u, weight, du, workspace, B, H, W, Ch, tx,  \
                                                          tx * ty, uwr_round_outputs());                          \
    } while (0)
    if (u_half) {
        DW_BWD(false, true, true);
    } else if (plain) {
        if (fast) DW_BWD(true, true);
        else DW_BWD(true, false);
    } else {
        if (fast) DW_BWD(false, true);
        else DW_BWD(false, false);
    }
#undef DW_BWD
    UWR_CHECK_LAUNCH("dwconv_bwd_kernel");
    (void)uwr_launch_pdl(dwconv_param_reduce_kernel, dim3(uwr_cdiv(11 * Ch, 128)), dim3(128), 0, stream, workspace, dweight, dbias, du_colsum, P, Ch);
    UWR_CHECK_LAUNCH("dwconv_param_reduce_kernel");
    return 0;
}
