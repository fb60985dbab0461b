// Adaptive sparse 8x8 window attention (WindowAttention_sparse.forward, AST.py:187-219) with the
// cyclic shift, window partition/reverse (AST.py:377-402,596-618) and the shift mask
// (AST.py:568-588) folded into address arithmetic.
//
//   S = scale*q k^T + table[idx(i,j),h] + mask ;  P = w0*softmax(S) + w1*relu(S)^2 ;  O = P v
//
// One (batch, window, head) tile = 64 tokens x HD.  QK^T / PV / the five backward products run
// on TF32 tensor cores (m16n8k8, fp32 accumulate); S, P, dS never leave registers / shared
// memory.  Algorithmic HBM bytes per tile: fwd 4*64*HD*4 (q,k,v in, o out),
// bwd 7*64*HD*4 (q,k,v,do in; dq,dk,dv out).
#include "uwr_common.cuh"
#include "uwr_tma.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int WIN = 8;
constexpr int NTOK = 64;
constexpr int ATT_THREADS = 128;  // 4 warps x 16 query rows
constexpr int PS_STRIDE = 72;     // P / dS staging, conflict-free for transposed A fragments
constexpr int NBINS = 225;

struct AttnParams {
    const float* q;
    long long ld_q;
    int q_off;
    const float* kv;
    long long ld_kv;
    int k_off, v_off;
    const float* table;
    const float* w_param;
    int B, H, W, heads, shift;
    float scale;
    int nWx, nW;  // windows per row / per image
    int rnd;      // round outputs to TF32 (they are GEMM operands)
    float* dq_colsum;   // backward only: column sums of dq / of dk, dv (bias gradients of the projections), or null
    float* dkv_colsum;
    int x3;       // 1: full-fp32 operands, error-compensated 3xTF32 on the cancelling products; 0: operands are exact TF32
                  //    values (uwr_attn_desc.operands_rounded), one pass is exact
};

__device__ __forceinline__ long long token_row(const AttnParams& p, int b, int wy, int wx, int n) {
    const int i = n >> 3, j = n & 7;
    int y = wy * WIN + i + p.shift;
    int x = wx * WIN + j + p.shift;
    if (y >= p.H) y -= p.H;
    if (x >= p.W) x -= p.W;
    return ((long long)b * p.H + y) * p.W + x;
}

__device__ __forceinline__ int region_code(const AttnParams& p, int wy, int wx, int n) {
    // region id of the token on the SHIFTED grid (AST.py:570-581)
    const int sy = wy * WIN + (n >> 3), sx = wx * WIN + (n & 7);
    const int ry = (sy >= p.H - WIN) + (sy >= p.H - p.shift);
    const int rx = (sx >= p.W - WIN) + (sx >= p.W - p.shift);
    return ry * 3 + rx;
}

__device__ __forceinline__ int bias_index(int i, int j) {
    return ((i >> 3) - (j >> 3) + WIN - 1) * (2 * WIN - 1) + ((i & 7) - (j & 7) + WIN - 1);
}

// stage 64 x HD tiles (fp32) into shared memory [64][HD+4].  All global loads of a batch are issued
// before the first shared-memory store (the one-load-at-a-time version left the kernel stalled on
// long_scoreboard with ~12% of the warps active).
template <int HD>
struct TileSrc {
    const float* base;
    long long ld;
    int col;
    float mul;
    float* dst;
    int rnd = 0;   // round the staged values to TF32 (exact-operand mode: scale * q is the only inexact tile)
};

template <int HD, int NT>
__device__ __forceinline__ void stage_tiles(const TileSrc<HD> (&t)[NT], const long long* rows) {
    constexpr int ST = HD + 4;
    constexpr int V4 = HD / 4;
    constexpr int PER = NTOK * V4 / ATT_THREADS;  // float4 per thread per tile
    constexpr int GROUP = (PER * NT <= 16) ? NT : 1;  // tiles whose loads are in flight together
#pragma unroll
    for (int g0 = 0; g0 < NT; g0 += GROUP) {
        float4 buf[GROUP][PER];
#pragma unroll
        for (int g = 0; g < GROUP; ++g)
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = threadIdx.x + i * ATT_THREADS;
                const int r = idx / V4, c4 = (idx % V4) * 4;
                buf[g][i] = *reinterpret_cast<const float4*>(t[g0 + g].base + rows[r] * t[g0 + g].ld + t[g0 + g].col + c4);
            }
#pragma unroll
        for (int g = 0; g < GROUP; ++g)
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int idx = threadIdx.x + i * ATT_THREADS;
                const int r = idx / V4, c4 = (idx % V4) * 4;
                const float m = t[g0 + g].mul;
                // full fp32 is kept in shared memory: S, dP, dQ, dK use error-compensated 3xTF32 (hi/lo
                // split at fragment load) because softmax' and the dS cancellations amplify rounding
                float4 v = make_float4(buf[g][i].x * m, buf[g][i].y * m, buf[g][i].z * m, buf[g][i].w * m);
                if (t[g0 + g].rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                *reinterpret_cast<float4*>(t[g0 + g].dst + r * ST + c4) = v;
            }
    }
}

// ldmatrix.x4 of 32-bit data viewed as pairs of .b16: four 8-row x 16-byte matrices, lanes 8m .. 8m+7 give the row
// addresses of matrix m, lane l receives the 32-bit word (l % 4) of row (l / 4) of each matrix
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr)
                 : "memory");
}

// 3xTF32 operand split.  The tensor core reads only the upper 19 bits of an fp32 operand (it
// truncates), so the "hi" part is the raw value itself and "lo" is what the truncation drops:
// lo = x - trunc(x), exact in fp32 (13 significant bits; the hardware truncating it again loses
// < 2^-20 |x|).  Two instructions per element instead of two roundings and a subtraction.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x);
    lo = __float_as_uint(x - __uint_as_float(hi & 0xFFFFE000u));
}

// X3: 3xTF32 split; otherwise a single operand rounded to nearest (a raw value would be truncated,
// which biases every product)
template <bool X3>
__device__ __forceinline__ void prep_tf32(float x, uint32_t& hi, uint32_t& lo) {
    if (X3) split_tf32(x, hi, lo);
    else { hi = f2tf32(x); lo = 0u; }
}

// acc[nt][4] (16 rows x 64 cols) = A[r0.., :HD] * Bm[:, :HD]^T   (both row-major fp32 [64][HD+4]),
// 3xTF32: a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (fp32-level accuracy on the tensor cores)
// EXACT (with x3 = false): both operands already hold TF32 values -- no rounding at the fragment load
template <int HD, bool EXACT = false>
__device__ __forceinline__ void mma_rows_x_rowsT(float (&acc)[8][4], const float* A, const float* Bm, int r0,
                                                 int g, int t, bool x3 = true) {
    constexpr int ST = HD + 4;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
    // Fragment loads by ldmatrix: a row of an 8x8 .b16 matrix is 16 bytes = four fp32 words, and lane l receives word
    // (l % 4) of row (l / 4) -- exactly element (g, t) of the m16n8k8 TF32 fragments when both operands are stored
    // [row][k] (rows 16-byte aligned, stride HD + 4 floats: the eight row addresses of a matrix fall into distinct bank
    // groups).  One .x4 = the whole A fragment, or the B fragments of two column tiles: 5 shared-memory instructions per
    // k step instead of 20 scalar loads (the backward was limited by shared-memory instruction issue).
    const int lane = g * 4 + t, lm = lane >> 3, lr = lane & 7;
    const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(A + (r0 + (lm & 1) * 8 + lr) * ST + (lm >> 1) * 4);
    const uint32_t b_addr = (uint32_t)__cvta_generic_to_shared(Bm + ((lm >> 1) * 8 + lr) * ST + (lm & 1) * 4);
#pragma unroll
    for (int ks = 0; ks < HD / 8; ++ks) {
        uint32_t ah[4], al[4];
        {
            uint32_t r[4];
            ldsm_x4(r, a_addr + ks * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) split_tf32(__uint_as_float(r[c]), ah[c], al[c]);
        }
        // the three passes run over all eight column tiles before the next pass touches the same
        // accumulator: eight independent MMAs between dependent ones (back-to-back dependent MMAs
        // left the tensor pipe waiting on its own latency)
        uint32_t bh[8][2], bl[8][2];
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t r[4];
            ldsm_x4(r, b_addr + (np * 16 * ST + ks * 8) * 4);
            split_tf32(__uint_as_float(r[0]), bh[2 * np][0], bl[2 * np][0]);
            split_tf32(__uint_as_float(r[1]), bh[2 * np][1], bl[2 * np][1]);
            split_tf32(__uint_as_float(r[2]), bh[2 * np + 1][0], bl[2 * np + 1][0]);
            split_tf32(__uint_as_float(r[3]), bh[2 * np + 1][1], bl[2 * np + 1][1]);
        }
        if (!x3 && !EXACT) {  // single pass: operands rounded to nearest (a raw value would be truncated)
#pragma unroll
            for (int c = 0; c < 4; ++c) ah[c] = (ah[c] + 0x1000u) & 0xFFFFE000u;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                bh[nt][0] = (bh[nt][0] + 0x1000u) & 0xFFFFE000u;
                bh[nt][1] = (bh[nt][1] + 0x1000u) & 0xFFFFE000u;
            }
        }
        if (x3) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], al, bh[nt]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], ah, bl[nt]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma_tf32_16x8x8(acc[nt], ah, bh[nt]);
    }
}

// out[nt2][4] (16 rows x HD) = P(regs, 16 x 64 in C-fragment layout) * Bm[64][HD+4]
// key permutation inside each k8 block: slot t <-> key 2t, slot t+4 <-> key 2t+1.
template <int HD, bool X3, bool EXACT = false>
__device__ __forceinline__ void mma_regs_x_rows(float (&out)[HD / 8][4], const float (&P)[8][4], const float* Bm,
                                                int g, int t, bool x3 = true) {
    constexpr int ST = HD + 4;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[n][c] = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        uint32_t a[4], al[4];
        prep_tf32<X3>(P[kb][0], a[0], al[0]);
        prep_tf32<X3>(P[kb][2], a[1], al[1]);
        prep_tf32<X3>(P[kb][1], a[2], al[2]);
        prep_tf32<X3>(P[kb][3], a[3], al[3]);
        uint32_t b[HD / 8][2], bl[HD / 8][2];
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            prep_tf32<X3>(Bm[(kb * 8 + 2 * t) * ST + n * 8 + g], b[n][0], bl[n][0]);
            prep_tf32<X3>(Bm[(kb * 8 + 2 * t + 1) * ST + n * 8 + g], b[n][1], bl[n][1]);
        }
        if (X3 && !x3 && !EXACT) {
#pragma unroll
            for (int c = 0; c < 4; ++c) a[c] = (a[c] + 0x1000u) & 0xFFFFE000u;
#pragma unroll
            for (int n = 0; n < HD / 8; ++n) {
                b[n][0] = (b[n][0] + 0x1000u) & 0xFFFFE000u;
                b[n][1] = (b[n][1] + 0x1000u) & 0xFFFFE000u;
            }
        }
        if (X3 && x3) {
#pragma unroll
            for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(out[n], al, b[n]);
#pragma unroll
            for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(out[n], a, bl[n]);
        }
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(out[n], a, b[n]);
    }
}

// S -> S*1 + bias + mask (q pre-scaled at staging), then row softmax statistics
__device__ __forceinline__ void add_bias_mask(float (&s)[8][4], const float* tab, const int* reg, int r0, int g,
                                              int t, bool masked) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = r0 + g + (c >> 1) * 8;
            const int j = nt * 8 + 2 * t + (c & 1);
            float v = s[nt][c] + tab[bias_index(i, j)];
            if (masked && reg[i] != reg[j]) v += -100.0f;
            s[nt][c] = v;
        }
}

__device__ __forceinline__ void row_softmax(const float (&s)[8][4], float (&p0)[8][4]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float m = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) m = fmaxf(m, fmaxf(s[nt][half * 2], s[nt][half * 2 + 1]));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            // exp(x) = 2^(x log2 e) on MUFU.EX2 without the denormal fix-up code of __expf (arguments <= 0)
            const float e0 = ex2_ftz((s[nt][half * 2] - m) * 1.4426950408889634f);
            const float e1 = ex2_ftz((s[nt][half * 2 + 1] - m) * 1.4426950408889634f);
            p0[nt][half * 2] = e0;
            p0[nt][half * 2 + 1] = e1;
            sum += e0 + e1;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = __fdividef(1.0f, sum);  // 1 <= sum <= 64
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            p0[nt][half * 2] *= inv;
            p0[nt][half * 2 + 1] *= inv;
        }
    }
}

__device__ __forceinline__ void fusion_weights(const float* w_param, float& w0, float& w1) {
    if (w_param == nullptr) {
        w0 = 1.f;
        w1 = 0.f;
        return;
    }
    const float e0 = expf(w_param[0]), e1 = expf(w_param[1]);
    w0 = e0 / (e0 + e1);
    w1 = e1 / (e0 + e1);
}

// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnParams p, float* __restrict__ out,
                                                               long long ld_out) {
    uwr_pdl_enter();
    constexpr int ST = HD + 4;
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* Ks = Qs + NTOK * ST;
    float* Vs = Ks + NTOK * ST;
    float* tab = Vs + NTOK * ST;                             // 225 (+pad to 228)
    int* reg = reinterpret_cast<int*>(tab + 228);            // 64
    long long* rows = reinterpret_cast<long long*>(reg + NTOK);  // 64

    const int tile = blockIdx.x;
    const int h = tile % p.heads;
    const int wlin = (tile / p.heads) % p.nW;
    const int b = tile / (p.heads * p.nW);
    const int wy = wlin / p.nWx, wx = wlin % p.nWx;

    for (int i = threadIdx.x; i < NBINS; i += ATT_THREADS) tab[i] = p.table[i * p.heads + h];
    if (threadIdx.x < NTOK) {
        rows[threadIdx.x] = token_row(p, b, wy, wx, threadIdx.x);
        reg[threadIdx.x] = p.shift > 0 ? region_code(p, wy, wx, threadIdx.x) : 0;
    }
    __syncthreads();
    {
        const TileSrc<HD> src[3] = {{p.q, p.ld_q, p.q_off + h * HD, p.scale, Qs},
                                    {p.kv, p.ld_kv, p.k_off + h * HD, 1.0f, Ks},
                                    {p.kv, p.ld_kv, p.v_off + h * HD, 1.0f, Vs}};
        stage_tiles<HD, 3>(src, rows);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16;

    float s[8][4], p0[8][4];
    mma_rows_x_rowsT<HD>(s, Qs, Ks, r0, g, t, p.x3);
    add_bias_mask(s, tab, reg, r0, g, t, p.shift > 0);
    row_softmax(s, p0);
    float w0, w1;
    fusion_weights(p.w_param, w0, w1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float r = fmaxf(s[nt][c], 0.f);
            p0[nt][c] = w0 * p0[nt][c] + w1 * r * r;
        }
    float o[HD / 8][4];
    mma_regs_x_rows<HD, false>(o, p0, Vs, g, t);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float* orow = out + rows[r0 + g + half * 8] * ld_out + h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n)
            *reinterpret_cast<float2*>(orow + n * 8 + 2 * t) =
                p.rnd ? make_float2(tf32_round(o[n][half * 2]), tf32_round(o[n][half * 2 + 1]))
                      : make_float2(o[n][half * 2], o[n][half * 2 + 1]);
    }
}

// Column sums of dq | dk | dv (the bias gradients of the q / kv projections, which would otherwise cost one more pass over
// the 3C-wide gradient: uwr_colsum, 0.64 ms per training step at B = 16).  Each thread keeps running sums of ITS
// accumulator slots (rows g and g + 8 folded, columns n*8 + 2t + e) across all the tiles it walks -- one add per value and
// tile; the eight row groups and the four warps are folded once, after the tile loop (fixed order -> deterministic).
// The sums are of the unrounded values (the rounding of the stored gradient is zero-mean noise on top of them).
template <int HD>
__device__ __forceinline__ void colsum_add(float (&cs)[HD / 8][2], const float (&v)[HD / 8][4]) {
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
        cs[n][0] += v[n][0] + v[n][2];
        cs[n][1] += v[n][1] + v[n][3];
    }
}
// one row of per-CTA partial sums: 225 bias-table bins, the two fusion-weight sums (+1 pad), 3 x HD column sums
__host__ __device__ constexpr int bwd_part_stride(int hd) { return NBINS + 3 + 3 * hd; }

// ------------------------------------------------------------------------------------------
// Backward.  grid = (ctas_per_head, heads); each CTA walks the (batch, window) tiles of one head
// so that the relative-position-bias gradient accumulates in registers and is binned once.
// X3: full-fp32 operands (3xTF32 compensation) vs exact TF32 operands (one pass); a template parameter, not a branch on
// p.x3: the kernel body is fully unrolled and ncu attributes 11 % of its stalls to instruction fetch -- only one of the
// two variants of every product is compiled in.
template <int HD, bool X3>
__global__ void __launch_bounds__(ATT_THREADS, HD <= 32 ? 3 : 1) attn_bwd_kernel(const AttnParams p,
                                                                              const float* __restrict__ dout,
                                                                              long long ld_dout,
                                                                              float* __restrict__ dq_buf,
                                                                              float* __restrict__ dkv_buf,
                                                                              float* __restrict__ partials) {
    uwr_pdl_enter();
    constexpr int ST = HD + 4;
    // the transposed products dV = P^T dO and dK = dS^T q need P / dS in shared memory; K and V are dead
    // by then, so (for HD >= 32) the 64 x 72 staging buffer aliases their tiles
    constexpr bool ALIAS = (2 * NTOK * ST >= NTOK * PS_STRIDE);
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;
    float* dOs = Qs + NTOK * ST;
    float* Ks = dOs + NTOK * ST;
    float* Vs = Ks + NTOK * ST;
    float* PD = ALIAS ? Ks : Vs + NTOK * ST;
    float* dSacc = (ALIAS ? Vs + NTOK * ST : PD + NTOK * PS_STRIDE);  // running sum of dS (bias-table gradient)
    float* tab = dSacc + NTOK * PS_STRIDE;
    int* reg = reinterpret_cast<int*>(tab + 228);
    long long* rows = reinterpret_cast<long long*>(reg + NTOK);
    __shared__ float red[2][ATT_THREADS / 32];
    __shared__ float colacc[ATT_THREADS / 32][3 * HD];   // per warp: column sums of dq | dk | dv

    const int h = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16;
    const bool want_colsum = p.dq_colsum != nullptr;

    for (int i = threadIdx.x; i < NBINS; i += ATT_THREADS) tab[i] = p.table[i * p.heads + h];
    float csq[HD / 8][2], csk[HD / 8][2], csv[HD / 8][2];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) csq[n][0] = csq[n][1] = csk[n][0] = csk[n][1] = csv[n][0] = csv[n][1] = 0.f;
    float w0, w1;
    fusion_weights(p.w_param, w0, w1);

    // each thread owns fixed (i, j) slots of the dS accumulator, so plain read-modify-write is safe
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<float2*>(dSacc + (r0 + g + half * 8) * PS_STRIDE + nt * 8 + 2 * t) = make_float2(0.f, 0.f);
    float g1 = 0.f, g2 = 0.f;

    const int ntiles = p.B * p.nW;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / p.nW, wlin = tile % p.nW;
        const int wy = wlin / p.nWx, wx = wlin % p.nWx;
        __syncthreads();  // previous iteration's consumers are done with smem
        if (threadIdx.x < NTOK) {
            rows[threadIdx.x] = token_row(p, b, wy, wx, threadIdx.x);
            reg[threadIdx.x] = p.shift > 0 ? region_code(p, wy, wx, threadIdx.x) : 0;
        }
        __syncthreads();
        {
            // exact-operand mode (!X3): k, v, dout are TF32 values already; scale * q is rounded here, once, so that no
            // product of this kernel converts a fragment (a third of the instruction stream before)
            const TileSrc<HD> src[4] = {{p.q, p.ld_q, p.q_off + h * HD, p.scale, Qs, X3 ? 0 : 1},
                                        {p.kv, p.ld_kv, p.k_off + h * HD, 1.0f, Ks},
                                        {p.kv, p.ld_kv, p.v_off + h * HD, 1.0f, Vs},
                                        {dout, ld_dout, h * HD, 1.0f, dOs}};
            stage_tiles<HD, 4>(src, rows);
        }
        __syncthreads();

        float pm[8][4], dp[8][4];  // final mixture P and dP -> dS
        {
            float s[8][4], p0[8][4];
            mma_rows_x_rowsT<HD, !X3>(s, Qs, Ks, r0, g, t, X3);
            add_bias_mask(s, tab, reg, r0, g, t, p.shift > 0);
            row_softmax(s, p0);
            mma_rows_x_rowsT<HD, !X3>(dp, dOs, Vs, r0, g, t, X3);  // dP = dO V^T
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float rowdot = 0.f;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) rowdot += dp[nt][half * 2 + e] * p0[nt][half * 2 + e];
                rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
                rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
                const int i = r0 + g + half * 8;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float dv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = half * 2 + e;
                        const float r = fmaxf(s[nt][c], 0.f);
                        const float P0 = p0[nt][c], dP = dp[nt][c];
                        g1 += dP * P0;
                        g2 += dP * r * r;
                        pm[nt][c] = tf32_round(w0 * P0 + w1 * r * r);
                        const float ds = w0 * P0 * (dP - rowdot) + w1 * 2.f * r * dP;
                        dp[nt][c] = X3 ? ds : tf32_round(ds);  // dp now holds dS (full fp32 when dQ / dK are error-compensated)
                        dv[e] = ds;
                    }
                    float2* accp = reinterpret_cast<float2*>(dSacc + i * PS_STRIDE + nt * 8 + 2 * t);
                    const float2 old = *accp;
                    *accp = make_float2(old.x + dv[0], old.y + dv[1]);
                }
            }
        }
        // dQ = scale * dS K  (Qs holds scale*q, so the chain rule adds one more factor scale)
        {
            float dq[HD / 8][4];
            mma_regs_x_rows<HD, true, !X3>(dq, dp, Ks, g, t, X3);  // dS rows sum to ~0: needs 3xTF32
            if (want_colsum) colsum_add<HD>(csq, dq);   // scaled by p.scale when it is folded, after the tile loop
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float* drow = dq_buf + rows[r0 + g + half * 8] * p.ld_q + p.q_off + h * HD;
#pragma unroll
                for (int n = 0; n < HD / 8; ++n) {
                    const float a = dq[n][half * 2] * p.scale, bq = dq[n][half * 2 + 1] * p.scale;
                    *reinterpret_cast<float2*>(drow + n * 8 + 2 * t) =
                        p.rnd ? make_float2(tf32_round(a), tf32_round(bq)) : make_float2(a, bq);
                }
            }
        }
        if (ALIAS) __syncthreads();  // everyone is done with K and V before P overwrites them

        // dV[j,:] = sum_i P[i,j] dO[i,:]   then   dK[j,:] = sum_i dS[i,j] (scale q)[i,:]
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            if (which == 1) __syncthreads();  // dV readers are done with the staging buffer
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = r0 + g + half * 8;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<float2*>(PD + i * PS_STRIDE + nt * 8 + 2 * t) =
                        which == 0 ? make_float2(pm[nt][half * 2], pm[nt][half * 2 + 1])
                                   : make_float2(dp[nt][half * 2], dp[nt][half * 2 + 1]);
            }
            __syncthreads();
            const float* Bm = which == 0 ? dOs : Qs;
            float acc[HD / 8][4];
#pragma unroll
            for (int n = 0; n < HD / 8; ++n)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                uint32_t a[4], al[4];
                // which == 0: P was rounded to TF32 when it was formed (pm), dO is rounded here
                const float pa0 = PD[(kb * 8 + t) * PS_STRIDE + r0 + g], pa1 = PD[(kb * 8 + t) * PS_STRIDE + r0 + g + 8];
                const float pa2 = PD[(kb * 8 + t + 4) * PS_STRIDE + r0 + g];
                const float pa3 = PD[(kb * 8 + t + 4) * PS_STRIDE + r0 + g + 8];
                if (which == 1) {
                    split_tf32(pa0, a[0], al[0]); split_tf32(pa1, a[1], al[1]);
                    split_tf32(pa2, a[2], al[2]); split_tf32(pa3, a[3], al[3]);
                } else {
                    a[0] = __float_as_uint(pa0); a[1] = __float_as_uint(pa1);
                    a[2] = __float_as_uint(pa2); a[3] = __float_as_uint(pa3);
                }
                uint32_t bb[HD / 8][2], bl[HD / 8][2];
#pragma unroll
                for (int n = 0; n < HD / 8; ++n) {
                    if (which == 1) {
                        split_tf32(Bm[(kb * 8 + t) * ST + n * 8 + g], bb[n][0], bl[n][0]);
                        split_tf32(Bm[(kb * 8 + t + 4) * ST + n * 8 + g], bb[n][1], bl[n][1]);
                    } else if (X3) {
                        bb[n][0] = f2tf32(Bm[(kb * 8 + t) * ST + n * 8 + g]);
                        bb[n][1] = f2tf32(Bm[(kb * 8 + t + 4) * ST + n * 8 + g]);
                    } else {   // dO holds TF32 values
                        bb[n][0] = __float_as_uint(Bm[(kb * 8 + t) * ST + n * 8 + g]);
                        bb[n][1] = __float_as_uint(Bm[(kb * 8 + t + 4) * ST + n * 8 + g]);
                    }
                }
                if (false) {   // (exact-operand mode: dS was rounded when it was formed, scale * q when it was staged)
#pragma unroll
                    for (int c = 0; c < 4; ++c) a[c] = (a[c] + 0x1000u) & 0xFFFFE000u;
#pragma unroll
                    for (int n = 0; n < HD / 8; ++n) {
                        bb[n][0] = (bb[n][0] + 0x1000u) & 0xFFFFE000u;
                        bb[n][1] = (bb[n][1] + 0x1000u) & 0xFFFFE000u;
                    }
                }
                if (which == 1 && X3) {  // dK: error-compensated
#pragma unroll
                    for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(acc[n], al, bb[n]);
#pragma unroll
                    for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(acc[n], a, bl[n]);
                }
#pragma unroll
                for (int n = 0; n < HD / 8; ++n) mma_tf32_16x8x8(acc[n], a, bb[n]);
            }
            const int off = which == 0 ? p.v_off : p.k_off;
            if (want_colsum) {
                if (which == 0) colsum_add<HD>(csv, acc);
                else colsum_add<HD>(csk, acc);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float* drow = dkv_buf + rows[r0 + g + half * 8] * p.ld_kv + off + h * HD;
#pragma unroll
                for (int n = 0; n < HD / 8; ++n)
                    *reinterpret_cast<float2*>(drow + n * 8 + 2 * t) =
                        p.rnd ? make_float2(tf32_round(acc[n][half * 2]), tf32_round(acc[n][half * 2 + 1]))
                              : make_float2(acc[n][half * 2], acc[n][half * 2 + 1]);
            }
        }
    }

    if (want_colsum) {   // fold the eight row groups (lanes with equal t), then the g == 0 lanes publish the warp's sums
#pragma unroll
        for (int n = 0; n < HD / 8; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = csq[n][e] * p.scale, b = csk[n][e], c = csv[n][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                if (g == 0) {
                    colacc[warp][n * 8 + 2 * t + e] = a;
                    colacc[warp][HD + n * 8 + 2 * t + e] = b;
                    colacc[warp][2 * HD + n * 8 + 2 * t + e] = c;
                }
            }
    }
    // ---- bin the accumulated dS into the 225 relative-position slots (deterministic) ----
    __syncthreads();
    g1 = warp_sum(g1);
    g2 = warp_sum(g2);
    if (lane == 0) {
        red[0][warp] = g1;
        red[1][warp] = g2;
    }
    __syncthreads();
    float* part = partials + ((long long)h * gridDim.x + blockIdx.x) * bwd_part_stride(HD);
    if (want_colsum)
        for (int i = threadIdx.x; i < 3 * HD; i += ATT_THREADS)   // (the __syncthreads above ordered the warps' slices)
            part[NBINS + 3 + i] = colacc[0][i] + colacc[1][i] + colacc[2][i] + colacc[3][i];
    for (int bin = threadIdx.x; bin < NBINS; bin += ATT_THREADS) {
        const int dy = bin / 15 - 7, dx = bin % 15 - 7;
        float sum = 0.f;
        for (int yj = max(0, -dy); yj < min(8, 8 - dy); ++yj)
            for (int xj = max(0, -dx); xj < min(8, 8 - dx); ++xj) {
                const int j = yj * 8 + xj, i = (yj + dy) * 8 + (xj + dx);
                sum += dSacc[i * PS_STRIDE + j];
            }
        part[bin] = sum;
    }
    if (threadIdx.x == 0) {
        part[NBINS] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
        part[NBINS + 1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
}

// One WARP per output element: lanes stride over the CTAs' partials (all loads independent), then a fixed-order shuffle
// fold -- deterministic.  (One THREAD per element, the first version, walked up to 444 partials as a serial chain of L2
// latencies, and thread 0 walked all heads x CTAs twice for the mixing-weight gradient: 29 us per launch under ncu, ten
// launches per AST step, for a few KB of output.)
__global__ void __launch_bounds__(128) attn_param_reduce_kernel(const float* __restrict__ partials,
                                                                const float* __restrict__ w_param, float* __restrict__ dtable,
                                                                float* __restrict__ dw, int heads, int ctas_per_head, int hd,
                                                                float* __restrict__ dq_colsum, float* __restrict__ dkv_colsum,
                                                                int q_off, int k_off, int v_off) {
    uwr_pdl_enter();
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int ps = bwd_part_stride(hd);
    const int n_table = NBINS * heads, n_all = (NBINS + 3 * hd) * heads;
    if (idx < n_table) {
        const int bin = idx / heads, h = idx % heads;
        float s = 0.f;
        for (int c = lane; c < ctas_per_head; c += 32) s += partials[((long long)h * ctas_per_head + c) * ps + bin];
        s = warp_sum(s);
        if (lane == 0) dtable[idx] = s;
    } else if (idx < n_all) {
        if (dq_colsum == nullptr) return;
        // column sums of dq | dk | dv, written with the column addressing of dq_buf / dkv_buf
        const int j = idx - n_table;
        const int h = j / (3 * hd), r = j % (3 * hd), which = r / hd, c = r % hd;
        float s = 0.f;
        for (int k = lane; k < ctas_per_head; k += 32) s += partials[((long long)h * ctas_per_head + k) * ps + NBINS + 3 + r];
        s = warp_sum(s);
        if (lane == 0) {
            if (which == 0) dq_colsum[q_off + h * hd + c] = s;
            else dkv_colsum[(which == 1 ? k_off : v_off) + h * hd + c] = s;
        }
    } else if (idx == n_all && dw != nullptr) {
        float g1 = 0.f, g2 = 0.f;
        for (int i = lane; i < heads * ctas_per_head; i += 32) {
            g1 += partials[(long long)i * ps + NBINS];
            g2 += partials[(long long)i * ps + NBINS + 1];
        }
        g1 = warp_sum(g1);
        g2 = warp_sum(g2);
        if (lane == 0) {
            float w0 = 1.f, w1 = 0.f;
            if (w_param) {
                const float e0 = expf(w_param[0]), e1 = expf(w_param[1]);
                w0 = e0 / (e0 + e1);
                w1 = e1 / (e0 + e1);
            }
            const float mix = w0 * g1 + w1 * g2;
            dw[0] = w0 * (g1 - mix);
            dw[1] = w1 * (g2 - mix);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Forward on the 5th-generation tensor cores (head_dim 32, shift 0 or 4).
//
// One item = TWO windows of one head stacked to a 128-row tile, so that S = Q K^T is a single
// M=128, N=128, K=32 tcgen05.mma tile (the two off-diagonal 64x64 blocks are never read) and
// O = P V is two M=128, N=32, K=64 tiles, D_w = P V_w, of which window w's 64 rows are read.
// The next item's Q/K/V tiles are requested as soon as S sits in TMEM, i.e. they land under the
// softmax / P V / store of the current item.
//   * Q / K / V rows arrive by TMA: an 8x8 window, cyclically shifted by 0 or 4, is four 4x4-pixel
//     quadrants that never wrap, i.e. four boxes (32 ch, 4 px, 4 rows) of a 4-D (C, W, H, B) tensor map
//     on the qkv token matrix; tokens are stored quadrant-major (a fixed permutation, undone by the
//     bias / mask / output index arithmetic).  Q, K use SWIZZLE_128B (K-major operands), V uses
//     SWIZZLE_128B_ATOM_32B (MN-major B operand of P V).
//   * 3xTF32 for S: the tensor core truncates fp32 operands, so hi = the TMA tile as it is and
//     lo = x - trunc(x) is written by the CTA into a second tile: S = Qlo Khi + Qhi Klo + Qhi Khi.
//   * softmax + relu^2 + learned fusion: one thread per query row on its 64 TMEM columns
//     (tcgen05.ld 32x32b), no shuffles; P (TF32-rounded) goes back to shared memory as the K-major A
//     operand of the second MMA (it aliases the Q/K tiles, which are dead by then).
//   * fp32 accumulators in TMEM: S in columns 0..127, O in columns 128..159 (256 allocated, 2 CTAs/SM).
#ifdef T5A_TIMING
__device__ long long t5a_dbg[16];
#define T5A_STAMP(k)                                                     \
    do {                                                                 \
        if (threadIdx.x == 0 && blockIdx.x == 0 && item_no == 3) t5a_dbg[k] = clock64(); \
    } while (0)
#else
#define T5A_STAMP(k)
#endif
constexpr int T5A_THREADS = 256;          // two threads per query row (32 score columns each)
constexpr int T5A_TILE = 128 * 32 * 4;    // one 128-row x 32-float operand tile
constexpr int T5A_MAXH = 16;              // heads whose bias tables are cached per CTA (C <= 512 at head_dim 32)
// dynamic shared memory: 6 operand tiles + bias tables of `heads` heads + regions + row max/sum exchange + barriers
// (+ 1 KB alignment slack): 108.6 KB at 8 heads -> two CTAs per SM (the 232 448-byte limit is tight: the table
// region is sized by the actual head count)
__host__ __device__ constexpr int t5a_tab_bytes(int heads) { return ((heads * 228 * 4 + 127) / 128) * 128; }
__host__ __device__ constexpr int t5a_smem_bytes(int heads) {
    return 6 * T5A_TILE + t5a_tab_bytes(heads) + 512 + 2048 + 64 + 1024;
}

__device__ __forceinline__ void t5a_token(int rr, int& i, int& j) {  // quadrant-major slot -> (row, col) in the window
    const int quad = rr >> 4;
    i = 4 * (quad >> 1) + ((rr >> 2) & 3);
    j = 4 * (quad & 1) + (rr & 3);
}

// the 24 quadrant boxes (2 windows x 4 quadrants x {Q, K, V}) of one item, one per lane of warp 0
__device__ __forceinline__ void t5a_issue_loads(const AttnParams& p, const CUtensorMap* mapQ, const CUtensorMap* mapK,
                                                const CUtensorMap* mapV, uint8_t* Qhi, uint8_t* Khi, uint8_t* Vs,
                                                uint64_t* bar, int item, int pairs, int lane) {
    using namespace uwr_tma;
    if (lane == 0) mbar_expect_tx(bar, 3 * T5A_TILE);
    __syncwarp();
    if (lane < 24) {
        const int h = item % p.heads;
        const int pr = (item / p.heads) % pairs;
        const int b = item / (p.heads * pairs);
        const int which = lane % 3, quad = (lane / 3) & 3, w = lane / 12;
        const int wlin = 2 * pr + w, wy = wlin / p.nWx, wx = wlin - wy * p.nWx;
        int y0 = wy * WIN + p.shift + 4 * (quad >> 1), x0 = wx * WIN + p.shift + 4 * (quad & 1);
        if (y0 >= p.H) y0 -= p.H;
        if (x0 >= p.W) x0 -= p.W;
        const int off = (w * 64 + quad * 16) * 128;
        if (which == 0) tma_load_4d(Qhi + off, mapQ, bar, p.q_off + h * 32, x0, y0, b);
        else if (which == 1) tma_load_4d(Khi + off, mapK, bar, p.k_off + h * 32, x0, y0, b);
        else tma_load_4d(Vs + off, mapV, bar, p.v_off + h * 32, x0, y0, b);
    }
}

__global__ void __launch_bounds__(T5A_THREADS, 2)
attn_fwd_t5_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const AttnParams p, float* __restrict__ out,
                   long long ld_out) {
    uwr_pdl_enter();
    using namespace uwr_tma;
    extern __shared__ uint8_t t5a_raw[];
    uint8_t* smem = t5a_raw + ((1024u - (smem_u32(t5a_raw) & 1023u)) & 1023u);
    uint8_t* Qhi = smem;
    uint8_t* Khi = smem + T5A_TILE;
    uint8_t* Qlo = smem + 2 * T5A_TILE;
    uint8_t* Klo = smem + 3 * T5A_TILE;
    uint8_t* Pm = Qlo;                        // compact P: [2 chunks of 32 K-columns][128 rows][128 B], aliases Qlo|Klo
    uint8_t* Vbuf = smem + 4 * T5A_TILE;      // two V tiles (the next item's V lands while this one's P V runs)
    float* tabs = reinterpret_cast<float*>(smem + 6 * T5A_TILE);   // [heads][228] relative-position bias tables
    const int tab_bytes = t5a_tab_bytes(p.heads);
    int* reg = reinterpret_cast<int*>(smem + 6 * T5A_TILE + tab_bytes);
    float* xch = reinterpret_cast<float*>(smem + 6 * T5A_TILE + tab_bytes + 512);   // [2 (max | sum)][2 halves][128 rows]
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 6 * T5A_TILE + tab_bytes + 512 + 2048);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_mma, 1);
        mbar_init_fence();
        tma_prefetch_map(&mapQ);
        tma_prefetch_map(&mapK);
        tma_prefetch_map(&mapV);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < p.heads * NBINS; i += T5A_THREADS) {   // table is (225, heads) in global memory
        const int bin = i / p.heads, hh = i - bin * p.heads;
        tabs[hh * 228 + bin] = p.table[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes

    // instruction descriptors: D = f32, A = B = tf32, N >> 3 at bit 17, M >> 4 at bit 24; bit 16 = B is MN-major
    constexpr uint32_t IDESC_S = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t IDESC_O = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

    float w0, w1;
    fusion_weights(p.w_param, w0, w1);
    const int half = tid >> 7, row = tid & 127;   // half: which 32 of the row's 64 score columns / 16 of its 32 outputs
    const int win = row >> 6, rr = row & 63;
    int ti, tj;
    t5a_token(rr, ti, tj);
    const int bias_base = (ti + WIN - 1) * (2 * WIN - 1) + tj + WIN - 1;
    const int n_nat = ti * 8 + tj;  // natural token index of this row inside its window

    const int pairs = p.nW >> 1;
    const int items = p.B * pairs * p.heads;
    uint32_t ph_load = 0, ph_mma = 0;
    if (warp == 0 && (int)blockIdx.x < items)
        t5a_issue_loads(p, &mapQ, &mapK, &mapV, Qhi, Khi, Vbuf, bar_load, blockIdx.x, pairs, lane);
    int vb = 0;
    int item_no = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, vb ^= 1, ++item_no) {
        T5A_STAMP(0);
        const int h = item % p.heads;
        const int pr = (item / p.heads) % pairs;
        const int b = item / (p.heads * pairs);
        uint8_t* Vs = Vbuf + vb * T5A_TILE;
        const float* tab = tabs + h * 228;
        const int wlin = 2 * pr + win, wy = wlin / p.nWx, wx = wlin - wy * p.nWx;
        const int myreg = p.shift > 0 ? region_code(p, wy, wx, n_nat) : 0;
        if (half == 0) reg[row] = myreg;   // the previous item's readers are past their last barrier
        const bool masked = p.shift > 0 && (wy == p.H / WIN - 1 || wx == p.nWx - 1);
        const long long orow = token_row(p, b, wy, wx, n_nat);

        mbar_wait(bar_load, ph_load);
        ph_load ^= 1;
        T5A_STAMP(1);
        // lo = x - trunc(x) for Q and K (3xTF32), V rounded to nearest in place (single-pass P V); nothing to do when
        // the operands arrive as exact TF32 values
        if (p.x3) {
            const float4* q4 = reinterpret_cast<const float4*>(Qhi);
            const float4* k4 = reinterpret_cast<const float4*>(Khi);
            float4* ql = reinterpret_cast<float4*>(Qlo);
            float4* kl = reinterpret_cast<float4*>(Klo);
            float4* v4 = reinterpret_cast<float4*>(Vs);
#pragma unroll
            for (int i = 0; i < T5A_TILE / 16 / T5A_THREADS; ++i) {
                const int idx = tid + i * T5A_THREADS;
                const float4 a = q4[idx], c = k4[idx], v = v4[idx];
                ql[idx] = make_float4(a.x - __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u),
                                      a.y - __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u),
                                      a.z - __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u),
                                      a.w - __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u));
                kl[idx] = make_float4(c.x - __uint_as_float(__float_as_uint(c.x) & 0xFFFFE000u),
                                      c.y - __uint_as_float(__float_as_uint(c.y) & 0xFFFFE000u),
                                      c.z - __uint_as_float(__float_as_uint(c.z) & 0xFFFFE000u),
                                      c.w - __uint_as_float(__float_as_uint(c.w) & 0xFFFFE000u));
                v4[idx] = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        T5A_STAMP(2);
        if (tid == 0) {
            tc_fence_after();
            const uint32_t qh = smem_u32(Qhi), kh = smem_u32(Khi), qlw = smem_u32(Qlo), klw = smem_u32(Klo);
            const int pass0 = p.x3 ? 0 : 2;   // exact TF32 operands: only the hi * hi pass
            for (int pass = pass0; pass < 3; ++pass) {
                const uint32_t a = pass == 0 ? qlw : qh, bb = pass == 1 ? klw : kh;
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                    umma_tf32(tmem_base, make_smem_desc(a + k8 * 32, 16, 1024, 2), make_smem_desc(bb + k8 * 32, 16, 1024, 2),
                              IDESC_S, (pass > pass0 || k8 > 0) ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        T5A_STAMP(3);
        // S is in TMEM: the Q/K tiles are dead (and the other V buffer since the previous item's P V) -> the next
        // item's tiles are requested now and land under this item's softmax / P V / store
        if (warp == 7 && item + (int)gridDim.x < items)
            t5a_issue_loads(p, &mapQ, &mapK, &mapV, Qhi, Khi, Vbuf + (vb ^ 1) * T5A_TILE, bar_load, item + gridDim.x, pairs,
                            lane);
        // ---- this thread's half of its query row: 32 scores (TMEM columns 64*win + 32*half .. +31)
        uint32_t sr[32];
        tmem_ld32_issue(trow + win * 64 + half * 32, sr);
        tmem_ld_wait();
        float sv[32];
        float mx = -INFINITY;
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) {
            // column token: quadrant-major slot m = 32*half + mm -> rows 4*half .. 4*half+3 of the window
            const int im = 4 * half + ((mm >> 2) & 3), jm = 4 * (mm >> 4) + (mm & 3);
            float v = __uint_as_float(sr[mm]) * p.scale + tab[bias_base - (im * (2 * WIN - 1) + jm)];
            if (masked && reg[win * 64 + half * 32 + mm] != myreg) v += -100.0f;
            sv[mm] = v;
            mx = fmaxf(mx, v);
        }
        // local softmax statistics of this half row, merged with the other half's in ONE exchange (online softmax)
        float sum = 0.f;
        float pe[32];
#pragma unroll
        for (int mm = 0; mm < 32; ++mm) {
            pe[mm] = ex2_ftz((sv[mm] - mx) * 1.4426950408889634f);
            sum += pe[mm];
        }
        reinterpret_cast<float2*>(xch)[half * 128 + row] = make_float2(mx, sum);
        __syncthreads();
        const float2 other = reinterpret_cast<const float2*>(xch)[(half ^ 1) * 128 + row];
        const float mall = fmaxf(mx, other.x);
        const float mine = ex2_ftz((mx - mall) * 1.4426950408889634f);
        const float tot = sum * mine + other.y * ex2_ftz((other.x - mall) * 1.4426950408889634f);
        const float inv = __fdividef(mine, tot) * w0;   // p_m = pe[m] * exp(mx - mall) / tot
        // ---- P = w0 softmax + w1 relu^2, TF32-rounded, into the compact K-major A tile (128B swizzle): row r
        // holds the 64 probabilities of ITS window (chunk `half` = its 32 columns); the product with the other
        // window's V is never read
        {
            const int sw = row & 7;
            uint8_t* prow = Pm + half * T5A_TILE + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int mm = c * 4 + q;
                    const float r = fmaxf(sv[mm], 0.f);
                    e[q] = tf32_round(pe[mm] * inv + w1 * r * r);
                }
                *reinterpret_cast<float4*>(prow + ((c ^ sw) << 4)) = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        T5A_STAMP(4);
        if (tid == 0) {
            tc_fence_after();
            const uint32_t pa = smem_u32(Pm), vbase = smem_u32(Vs);
#pragma unroll
            for (int w = 0; w < 2; ++w)  // D_w = P (128 x 64) V_w (64 x 32): rows of window w are the valid ones
#pragma unroll
                for (int k8 = 0; k8 < 8; ++k8)
                    umma_tf32(tmem_base + 128 + 32 * w, make_smem_desc(pa + (k8 >> 2) * T5A_TILE + (k8 & 3) * 32, 16, 1024, 2),
                              make_smem_desc(vbase + w * 8192 + k8 * 1024, 4096, 512, 1), IDESC_O, k8 > 0 ? 1u : 0u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        T5A_STAMP(5);
        uint32_t orr[16];
        tmem_ld16_issue(trow + 128 + 32 * win + 16 * half, orr);
        tmem_ld_wait();
        float* op = out + orow * ld_out + h * 32 + 16 * half;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 v = make_float4(__uint_as_float(orr[4 * c]), __uint_as_float(orr[4 * c + 1]),
                                   __uint_as_float(orr[4 * c + 2]), __uint_as_float(orr[4 * c + 3]));
            if (p.rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
            *reinterpret_cast<float4*>(op + 4 * c) = v;
        }
        tc_fence_before();
        __syncthreads();  // every warp has read its TMEM rows before the next item's MMAs overwrite them
        T5A_STAMP(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// uwr_set_attn_tcgen05: 0 = never, 1 = whenever the shape is eligible, 2 = auto (default): when the operands arrive as
// exact TF32 values (uwr_attn_desc.operands_rounded — the training / inference path of the models).  Measured on B200
// (32 768 tiles, head_dim 32): with full-fp32 operands the tcgen05 kernel needs the lo split and three S passes and
// takes 0.347 ms against 0.316 ms for mma.sync; with exact TF32 operands it is one pass, no split: 0.279 ms against
// 0.285 ms.  Each item is still one serial chain TMA -> MMA -> softmax -> MMA -> store with two items per SM.
int g_attn_t5 = 2;

int launch_fwd_t5(const AttnParams& p, float* out, long long ld_out, cudaStream_t stream) {
    CUtensorMap mq, mk, mv;
    const int box[4] = {32, 4, 4, 1};
    {
        const long long dims[4] = {p.ld_q, p.W, p.H, p.B};
        const long long str[3] = {p.ld_q, (long long)p.W * p.ld_q, (long long)p.H * p.W * p.ld_q};
        if (uwr_tma::encode_f32(&mq, p.q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    }
    {
        const long long dims[4] = {p.ld_kv, p.W, p.H, p.B};
        const long long str[3] = {p.ld_kv, (long long)p.W * p.ld_kv, (long long)p.H * p.W * p.ld_kv};
        if (uwr_tma::encode_f32(&mk, p.kv, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
        if (uwr_tma::encode_f32(&mv, p.kv, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -3;
    }
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(attn_fwd_t5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      t5a_smem_bytes(T5A_MAXH)));
        configured = true;
    }
    const long long items = (long long)p.B * (p.nW / 2) * p.heads;
    const int grid = (int)(items < 2LL * uwr_sm_count() ? items : 2LL * uwr_sm_count());
    (void)uwr_launch_pdl(attn_fwd_t5_kernel, dim3(grid), dim3(T5A_THREADS), t5a_smem_bytes(p.heads), stream, mq, mk, mv, p, out, ld_out);
    UWR_CHECK_LAUNCH("attn_fwd_t5_kernel");
    return 0;
}

template <int HD>
constexpr int fwd_smem() { return (3 * NTOK * (HD + 4) + 228 + NTOK) * 4 + NTOK * 8; }
template <int HD>
constexpr int bwd_smem() {
    return (4 * NTOK * (HD + 4) + ((2 * (HD + 4) >= PS_STRIDE) ? 1 : 2) * NTOK * PS_STRIDE + 228 + NTOK) * 4 + NTOK * 8;
}

// CTAs of the backward kernel that are resident per SM (registers and shared memory of the compiled kernel, asked from
// the runtime once per head_dim; the smaller of the two operand-mode variants)
template <int HD>
int bwd_resident_per_sm() {
    static int n = 0;
    if (n == 0) {
        int a = 0, b = 0;
        cudaFuncSetAttribute(attn_bwd_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<HD>());
        cudaFuncSetAttribute(attn_bwd_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<HD>());
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, attn_bwd_kernel<HD, true>, ATT_THREADS, bwd_smem<HD>()) != cudaSuccess) a = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, attn_bwd_kernel<HD, false>, ATT_THREADS, bwd_smem<HD>()) != cudaSuccess) b = 1;
        n = a < b ? a : b;
        if (n < 1) n = 1;
        (void)cudaGetLastError();
    }
    return n;
}

// The backward grid is (CTAs per head, heads), every CTA strides over its head's windows: ONE resident wave.  The CTA
// count per head is rounded DOWN -- rounding up (the first version) put 56 x 8 = 448 CTAs on 444 slots for 8 heads, and
// the four left-over CTAs ran their ~18 windows after everyone else had finished: up to twice the kernel time.
int bwd_ctas_per_head(const uwr_attn_desc* d) {
    const int tiles = d->B * (d->H / WIN) * (d->W / WIN);
    int res = 1;
    switch (d->head_dim) {
        case 8: res = bwd_resident_per_sm<8>(); break;
        case 16: res = bwd_resident_per_sm<16>(); break;
        case 32: res = bwd_resident_per_sm<32>(); break;
        case 64: res = bwd_resident_per_sm<64>(); break;
        case 128: res = bwd_resident_per_sm<128>(); break;
    }
    int c = (res * uwr_sm_count()) / d->heads;
    if (c > tiles) c = tiles;
    if (c < 1) c = 1;
    return c;
}

int fill_params(const uwr_attn_desc* d, AttnParams& p, const char* who) {
    UWR_REQUIRE(d && d->q && d->kv && d->bias_table, "%s: null pointer", who);
    UWR_REQUIRE(d->H % WIN == 0 && d->W % WIN == 0 && d->H >= WIN && d->W >= WIN, "%s: H,W must be multiples of 8", who);
    UWR_REQUIRE(d->shift >= 0 && d->shift < WIN, "%s: shift must be in [0,8)", who);
    UWR_REQUIRE(d->ld_q % 4 == 0 && d->ld_kv % 4 == 0 && d->q_off % 4 == 0 && d->k_off % 4 == 0 && d->v_off % 4 == 0,
                "%s: leading dims / offsets must be multiples of 4", who);
    p.q = d->q; p.ld_q = d->ld_q; p.q_off = d->q_off;
    p.kv = d->kv; p.ld_kv = d->ld_kv; p.k_off = d->k_off; p.v_off = d->v_off;
    p.table = d->bias_table; p.w_param = d->w_param;
    p.B = d->B; p.H = d->H; p.W = d->W; p.heads = d->heads; p.shift = d->shift; p.scale = d->scale;
    p.nWx = d->W / WIN; p.nW = (d->H / WIN) * (d->W / WIN);
    p.rnd = uwr_round_outputs();
    p.x3 = d->operands_rounded ? 0 : 1;
    p.dq_colsum = d->dq_colsum;
    p.dkv_colsum = d->dkv_colsum;
    return 0;
}

template <int HD>
int launch_fwd(const AttnParams& p, float* out, long long ld_out, cudaStream_t stream) {
    auto kern = attn_fwd_kernel<HD>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<HD>()));
        configured = true;
    }
    const long long tiles = (long long)p.B * p.nW * p.heads;
    (void)uwr_launch_pdl(kern, dim3((unsigned)tiles), dim3(ATT_THREADS), fwd_smem<HD>(), stream, p, out, ld_out);
    UWR_CHECK_LAUNCH("attn_fwd_kernel");
    return 0;
}

template <int HD, bool X3>
int launch_bwd_x(const AttnParams& p, const float* dout, long long ld_dout, float* dq, float* dkv, float* partials,
                 int cph, cudaStream_t stream) {
    auto kern = attn_bwd_kernel<HD, X3>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<HD>()));
        configured = true;
    }
    (void)uwr_launch_pdl(kern, dim3(dim3(cph, p.heads)), dim3(ATT_THREADS), bwd_smem<HD>(), stream, p, dout, ld_dout, dq, dkv, partials);
    UWR_CHECK_LAUNCH("attn_bwd_kernel");
    return 0;
}
template <int HD>
int launch_bwd(const AttnParams& p, const float* dout, long long ld_dout, float* dq, float* dkv, float* partials,
               int cph, cudaStream_t stream) {
    return p.x3 ? launch_bwd_x<HD, true>(p, dout, ld_dout, dq, dkv, partials, cph, stream)
                : launch_bwd_x<HD, false>(p, dout, ld_dout, dq, dkv, partials, cph, stream);
}

}  // namespace

extern "C" int uwr_window_attn_fwd(const uwr_attn_desc* d, float* out, long long ld_out, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AttnParams p;
    if (int rc = fill_params(d, p, "uwr_window_attn_fwd")) return rc;
    UWR_REQUIRE(out && ld_out % 2 == 0, "uwr_window_attn_fwd: bad output");
    // tcgen05 / TMA path: two windows per 128-row MMA tile
    const bool want_t5 = g_attn_t5 == 1 || (g_attn_t5 == 2 && d->operands_rounded);
    if (want_t5 && d->head_dim == 32 && d->heads <= T5A_MAXH && (d->shift == 0 || d->shift == 4) && p.nW % 2 == 0 &&
        ld_out % 4 == 0 &&
        (((uintptr_t)d->q | (uintptr_t)d->kv | (uintptr_t)out) & 15) == 0)
        return launch_fwd_t5(p, out, ld_out, stream);
    switch (d->head_dim) {
        case 8: return launch_fwd<8>(p, out, ld_out, stream);
        case 16: return launch_fwd<16>(p, out, ld_out, stream);
        case 32: return launch_fwd<32>(p, out, ld_out, stream);
        case 64: return launch_fwd<64>(p, out, ld_out, stream);
        case 128: return launch_fwd<128>(p, out, ld_out, stream);
    }
    uwr_set_error("uwr_window_attn_fwd: head_dim %d unsupported (8,16,32,64,128)", d->head_dim);
    return -1;
}

#ifdef T5A_TIMING
extern "C" int uwr_attn_t5_debug(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, t5a_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int uwr_set_attn_tcgen05(int mode) {
    g_attn_t5 = (mode == 2) ? 2 : (mode ? 1 : 0);
    return 0;
}

extern "C" size_t uwr_window_attn_bwd_workspace_bytes(const uwr_attn_desc* d) {
    return (size_t)d->heads * bwd_ctas_per_head(d) * bwd_part_stride(d->head_dim) * sizeof(float);
}

extern "C" int uwr_window_attn_bwd(const uwr_attn_desc* d, const float* dout, long long ld_dout, float* dq_buf,
                                   float* dkv_buf, float* dbias_table, float* dw, float* workspace,
                                   uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AttnParams p;
    if (int rc = fill_params(d, p, "uwr_window_attn_bwd")) return rc;
    UWR_REQUIRE(dout && dq_buf && dkv_buf && dbias_table && workspace && ld_dout % 4 == 0, "uwr_window_attn_bwd: bad args");
    const int cph = bwd_ctas_per_head(d);
    int rc;
    switch (d->head_dim) {
        case 8: rc = launch_bwd<8>(p, dout, ld_dout, dq_buf, dkv_buf, workspace, cph, stream); break;
        case 16: rc = launch_bwd<16>(p, dout, ld_dout, dq_buf, dkv_buf, workspace, cph, stream); break;
        case 32: rc = launch_bwd<32>(p, dout, ld_dout, dq_buf, dkv_buf, workspace, cph, stream); break;
        case 64: rc = launch_bwd<64>(p, dout, ld_dout, dq_buf, dkv_buf, workspace, cph, stream); break;
        case 128: rc = launch_bwd<128>(p, dout, ld_dout, dq_buf, dkv_buf, workspace, cph, stream); break;
        default:
            uwr_set_error("uwr_window_attn_bwd: head_dim %d unsupported (8,16,32,64,128)", d->head_dim);
            return -1;
    }
    if (rc) return rc;
    UWR_REQUIRE((d->dq_colsum == nullptr) == (d->dkv_colsum == nullptr), "uwr_window_attn_bwd: dq_colsum and dkv_colsum go together");
    // one warp per output element (+ one for the mixing-weight gradient), four warps per CTA
    (void)uwr_launch_pdl(attn_param_reduce_kernel, dim3(uwr_cdiv((NBINS + 3 * d->head_dim) * d->heads + 1, 4)), dim3(128), 0, stream, 
        workspace, d->w_param, dbias_table, dw, d->heads, cph, d->head_dim, d->dq_colsum, d->dkv_colsum, d->q_off, d->k_off,
        d->v_off);
    UWR_CHECK_LAUNCH("attn_param_reduce_kernel");
    return 0;
}
