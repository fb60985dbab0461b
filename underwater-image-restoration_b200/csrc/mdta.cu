// MDTA channel attention (SpectralTransformer.py:92-113) as two batched, per-head bandwidth kernels:
//
//   uwr_mdta_gram :  G[b,h,i,j] = sum_l X[b,l,h*c+i] * Y[b,l,h*c+j]      (q^T k over the H*W tokens,
//                    sqx[b,ch]  = sum_l X[b,l,ch]^2,  sqy likewise         line 100; the L2 norms of line 99)
//   uwr_mdta_apply:  out[b,l,h*c+i] = sum_j M'[b,h,i,j] * X[b,l,h*c+j]    (attn @ v, lines 101,109,113)
//                                    [+ diag[b,h*c+i] * Yd[b,l,h*c+i]]    (the norm terms of the backward)
//
// Tokens are rows of (B*L, ld) matrices (NHWC), channels of a head are contiguous, c = C/heads in
// {8,16,32,64}; both kernels read every token once (algorithmic bytes: gram 2*L*C*4 per image,
// apply 2*L*C*4 [+ L*C*4 with the diagonal term]) -> HBM-bound, AI = c/4 flop/B, far below a tcgen05
// atom (M = c <= 64 output rows), so the products run on m16n8k8 TF32 mma.sync with fp32 accumulate.
// The backward of the pair is the same two kernels: dA = gram(dout, v), dv = apply(dout, A^T),
// dq = apply(k, dG) + 2 dsq_q q, dk = apply(q, dG^T) + 2 dsq_k k.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int MD_THREADS = 128;
constexpr int GR_TOK = 32;   // tokens per staged tile (gram)
constexpr int AP_TOK = 64;   // tokens per tile (apply): 16 per warp

// ------------------------------------------------------------------------------------------ gram
// grid = (chunks, B).  Each CTA reduces a contiguous token range of one image; partial results go to
// workspace[b][chunk][heads*c*c + 2*C] and are summed by mdta_gram_reduce_kernel (deterministic).
template <int CH>   // CH = c
__global__ void __launch_bounds__(MD_THREADS)
mdta_gram_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ Y, long long ldy, int L,
                 int heads, int tiles_per_cta, float* __restrict__ partials, int want_sq) {
    uwr_pdl_enter();
    const int C = heads * CH;
    const int ST = C + 8;   // t*ST + g hits 32 distinct banks for the transposed fragment reads
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                       // [2][GR_TOK][ST]
    float* Ys = Xs + 2 * GR_TOK * ST;       // [2][GR_TOK][ST]

    const int b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ntiles = (L + GR_TOK - 1) / GR_TOK;
    const int t_begin = chunk * tiles_per_cta;
    const int t_end = min(ntiles, t_begin + tiles_per_cta);
    const float* Xb = X + (long long)b * L * ldx;
    const float* Yb = Y + (long long)b * L * ldy;

    // output tiles of 16 (i) x 8 (j): heads * (CH/16) * (CH/8), dealt round-robin to the 4 warps
    constexpr int MT = CH / 16 > 0 ? CH / 16 : 1, NT = CH / 8;
    constexpr int MAX_PER_WARP = 8;    // heads * c * c <= 4096 (checked by the launcher)
    const int out_tiles = heads * MT * NT;
    float acc[MAX_PER_WARP][4];
#pragma unroll
    for (int i = 0; i < MAX_PER_WARP; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    float sq0 = 0.f, sq1 = 0.f, sq2 = 0.f, sq3 = 0.f;   // squared norms of columns tid + 128*s of [X | Y]

    auto issue = [&](int tile, int buf) {
        const int v4 = C / 4;
        for (int idx = tid; idx < GR_TOK * v4; idx += MD_THREADS) {
            const int r = idx / v4, c4 = (idx - r * v4) * 4;
            const long long row = (long long)tile * GR_TOK + r;
            const bool ok = row < L;   // rows past the image are zero-filled (src-size 0)
            cp_async16(Xs + (buf * GR_TOK + r) * ST + c4, Xb + (ok ? row : 0) * ldx + c4, ok);
            cp_async16(Ys + (buf * GR_TOK + r) * ST + c4, Yb + (ok ? row : 0) * ldy + c4, ok);
        }
        cp_async_commit();
    };

    if (t_begin < t_end) issue(t_begin, 0);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int buf = (tile - t_begin) & 1;
        if (tile + 1 < t_end) {
            issue(tile + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* xs = Xs + buf * GR_TOK * ST;
        const float* ys = Ys + buf * GR_TOK * ST;
#pragma unroll
        for (int w = 0; w < MAX_PER_WARP; ++w) {
            const int ot = warp + 4 * w;
            if (ot >= out_tiles) break;
            const int h = ot / (MT * NT), rem = ot - h * (MT * NT);
            const int i0 = h * CH + (rem / NT) * 16, j0 = h * CH + (rem % NT) * 8;
            const bool hi_rows = (CH >= 16);   // c = 8: rows 8..15 of the M tile do not exist
#pragma unroll
            for (int k0 = 0; k0 < GR_TOK; k0 += 8) {
                uint32_t a[4], bb[2];
                a[0] = f2tf32(xs[(k0 + t) * ST + i0 + g]);
                a[2] = f2tf32(xs[(k0 + t + 4) * ST + i0 + g]);
                a[1] = hi_rows ? f2tf32(xs[(k0 + t) * ST + i0 + g + 8]) : 0u;
                a[3] = hi_rows ? f2tf32(xs[(k0 + t + 4) * ST + i0 + g + 8]) : 0u;
                bb[0] = f2tf32(ys[(k0 + t) * ST + j0 + g]);
                bb[1] = f2tf32(ys[(k0 + t + 4) * ST + j0 + g]);
                mma_tf32_16x8x8(acc[w], a, bb);
            }
        }
        if (want_sq) {   // squared column norms in full fp32 (F.normalize, line 99)
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {   // 2*C <= 512 columns, 128 threads
                const int ch = tid + s2 * MD_THREADS;
                if (ch < 2 * C) {
                    const float* src = ch < C ? xs + ch : ys + (ch - C);
                    float s = 0.f;
#pragma unroll 8
                    for (int r = 0; r < GR_TOK; ++r) {
                        const float v = src[r * ST];
                        s = fmaf(v, v, s);
                    }
                    if (s2 == 0) sq0 += s; else if (s2 == 1) sq1 += s; else if (s2 == 2) sq2 += s; else sq3 += s;
                }
            }
        }
        __syncthreads();   // the other buffer is refilled by the next iteration's cp.async
    }

    float* part = partials + ((long long)b * gridDim.x + chunk) * (heads * CH * CH + 2 * C);
#pragma unroll
    for (int w = 0; w < MAX_PER_WARP; ++w) {
        const int ot = warp + 4 * w;
        if (ot >= out_tiles) break;
        const int h = ot / (MT * NT), rem = ot - h * (MT * NT);
        const int i0 = (rem / NT) * 16, j0 = (rem % NT) * 8;
        float* gh = part + h * CH * CH;
        *reinterpret_cast<float2*>(gh + (i0 + g) * CH + j0 + 2 * t) = make_float2(acc[w][0], acc[w][1]);
        if (CH >= 16) *reinterpret_cast<float2*>(gh + (i0 + g + 8) * CH + j0 + 2 * t) = make_float2(acc[w][2], acc[w][3]);
    }
    if (want_sq) {
        const float sq[4] = {sq0, sq1, sq2, sq3};
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2)
            if (tid + s2 * MD_THREADS < 2 * C) part[heads * CH * CH + tid + s2 * MD_THREADS] = sq[s2];
    }
}

__global__ void mdta_gram_reduce_kernel(const float* __restrict__ partials, int chunks, int per, int gsz, int C,
                                        float* __restrict__ G, float* __restrict__ sqx, float* __restrict__ sqy) {
    uwr_pdl_enter();
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const float* p = partials + (long long)b * chunks * per + i;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += p[(long long)c * per];
    if (i < gsz) G[(long long)b * gsz + i] = s;
    else if (i < gsz + C) { if (sqx) sqx[(long long)b * C + (i - gsz)] = s; }
    else if (sqy) sqy[(long long)b * C + (i - gsz - C)] = s;
}

int gram_chunks(int B, int L) {
    const int tiles = (L + GR_TOK - 1) / GR_TOK;
    int chunks = (4 * uwr_sm_count() + B - 1) / B;
    if (chunks > tiles) chunks = tiles;
    if (chunks < 1) chunks = 1;
    return chunks;
}

// ----------------------------------------------------------------------------------------- apply
template <int CH>
__global__ void __launch_bounds__(MD_THREADS)
mdta_apply_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ Mx, int transpose,
                  const float* __restrict__ Yd, long long ldy, const float* __restrict__ diag, float* __restrict__ out,
                  long long ldo, int B, int L, int heads, int round_out) {
    uwr_pdl_enter();
    const int C = heads * CH;
    const int ST = C + 4;     // g*ST + t: conflict-free row-major fragment reads
    constexpr int MS = CH + 4;
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                     // [AP_TOK][ST]
    float* Ms = Xs + AP_TOK * ST;         // [heads][CH][MS]   Ms[h][i][j] = M'[i][j]
    float* Ds = Ms + heads * CH * MS;     // [C] diagonal

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_per_img = (L + AP_TOK - 1) / AP_TOK;
    const int total = B * tiles_per_img;
    int cur_b = -1;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int l0 = (tile - b * tiles_per_img) * AP_TOK;   // first token of the tile inside its image
        const long long row0 = (long long)b * L + l0;
        const int valid = min(AP_TOK, L - l0);
        __syncthreads();   // previous tile's readers are done with Xs (and Ms when the image changes)
        {
            const int v4 = C / 4;
            for (int idx = tid; idx < AP_TOK * v4; idx += MD_THREADS) {
                const int r = idx / v4, c4 = (idx - r * v4) * 4;
                cp_async16(Xs + r * ST + c4, X + (row0 + (r < valid ? r : 0)) * ldx + c4, r < valid);
            }
            cp_async_commit();
        }
        if (b != cur_b) {
            cur_b = b;
            const float* mb = Mx + (long long)b * heads * CH * CH;
            for (int idx = tid; idx < heads * CH * CH; idx += MD_THREADS) {
                const int h = idx / (CH * CH), r = (idx / CH) % CH, c = idx % CH;
                const float v = tf32_round(mb[idx]);
                if (transpose) Ms[(h * CH + c) * MS + r] = v;
                else Ms[(h * CH + r) * MS + c] = v;
            }
            if (diag != nullptr)
                for (int ch = tid; ch < C; ch += MD_THREADS) Ds[ch] = diag[(long long)b * C + ch];
        }
        cp_async_wait<0>();
        __syncthreads();
        const int r0 = warp * 16;
        for (int h = 0; h < heads; ++h) {
            float acc[CH / 8][4];
#pragma unroll
            for (int n = 0; n < CH / 8; ++n)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < CH; k0 += 8) {
                uint32_t a[4];
                a[0] = f2tf32(Xs[(r0 + g) * ST + h * CH + k0 + t]);
                a[1] = f2tf32(Xs[(r0 + g + 8) * ST + h * CH + k0 + t]);
                a[2] = f2tf32(Xs[(r0 + g) * ST + h * CH + k0 + t + 4]);
                a[3] = f2tf32(Xs[(r0 + g + 8) * ST + h * CH + k0 + t + 4]);
#pragma unroll
                for (int n = 0; n < CH / 8; ++n) {
                    uint32_t bb[2];   // B[k=j][n=i] = M'[i][j]
                    bb[0] = __float_as_uint(Ms[(h * CH + n * 8 + g) * MS + k0 + t]);
                    bb[1] = __float_as_uint(Ms[(h * CH + n * 8 + g) * MS + k0 + t + 4]);
                    mma_tf32_16x8x8(acc[n], a, bb);
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (r0 + g + half * 8 >= valid) continue;
                const long long row = row0 + r0 + g + half * 8;
#pragma unroll
                for (int n = 0; n < CH / 8; ++n) {
                    const int ch = h * CH + n * 8 + 2 * t;
                    float2 o = make_float2(acc[n][half * 2], acc[n][half * 2 + 1]);
                    if (diag != nullptr) {
                        const float2 y = *reinterpret_cast<const float2*>(Yd + row * ldy + ch);
                        o.x = fmaf(Ds[ch], y.x, o.x);
                        o.y = fmaf(Ds[ch + 1], y.y, o.y);
                    }
                    if (round_out) o = make_float2(tf32_round(o.x), tf32_round(o.y));
                    *reinterpret_cast<float2*>(out + row * ldo + ch) = o;
                }
            }
        }
    }
}

template <int CH>
int launch_gram(const float* X, long long ldx, const float* Y, long long ldy, int B, int L, int heads, float* G,
                float* sqx, float* sqy, float* ws, cudaStream_t stream) {
    const int C = heads * CH;
    const int chunks = gram_chunks(B, L);
    const int tiles = (L + GR_TOK - 1) / GR_TOK;
    const int tpc = uwr_cdiv(tiles, chunks);
    const int smem = 4 * GR_TOK * (C + 8) * 4;
    auto kern = mdta_gram_kernel<CH>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * GR_TOK * (256 + 8) * 4));
        configured = true;
    }
    (void)uwr_launch_pdl(kern, dim3(dim3(chunks, B)), dim3(MD_THREADS), smem, stream, X, ldx, Y, ldy, L, heads, tpc, ws, (sqx || sqy) ? 1 : 0);
    UWR_CHECK_LAUNCH("mdta_gram_kernel");
    const int gsz = heads * CH * CH, per = gsz + 2 * C;
    (void)uwr_launch_pdl(mdta_gram_reduce_kernel, dim3(dim3(uwr_cdiv(per, 128), B)), dim3(128), 0, stream, ws, chunks, per, gsz, C, G, sqx, sqy);
    UWR_CHECK_LAUNCH("mdta_gram_reduce_kernel");
    return 0;
}

template <int CH>
int launch_apply(const float* X, long long ldx, const float* Mx, int transpose, const float* Yd, long long ldy,
                 const float* diag, float* out, long long ldo, int B, int L, int heads, cudaStream_t stream) {
    const int C = heads * CH;
    const int smem = (AP_TOK * (C + 4) + heads * CH * (CH + 4) + C) * 4;
    auto kern = mdta_apply_kernel<CH>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (AP_TOK * (256 + 4) + 256 * (CH + 4) + 256) * 4));
        configured = true;
    }
    const long long total = (long long)B * ((L + AP_TOK - 1) / AP_TOK);
    int grid = 4 * uwr_sm_count();
    if (grid > total) grid = (int)total;
    (void)uwr_launch_pdl(kern, dim3(grid), dim3(MD_THREADS), smem, stream, X, ldx, Mx, transpose, Yd, ldy, diag, out, ldo, B, L, heads,
                                             uwr_round_outputs());
    UWR_CHECK_LAUNCH("mdta_apply_kernel");
    return 0;
}

int check_common(const char* who, int B, int L, int heads, int c) {
    UWR_REQUIRE(B >= 1 && L >= 1, "%s: B, L must be positive", who);
    UWR_REQUIRE(c == 8 || c == 16 || c == 32 || c == 64, "%s: channels per head %d unsupported (8,16,32,64)", who, c);
    UWR_REQUIRE(heads >= 1 && heads * c <= 256, "%s: heads * c must be <= 256", who);
    return 0;
}

}  // namespace

extern "C" size_t uwr_mdta_gram_workspace_bytes(int B, int L, int heads, int c) {
    if (B < 1 || L < 1) return 0;
    return (size_t)B * gram_chunks(B, L) * ((size_t)heads * c * c + 2 * (size_t)heads * c) * sizeof(float);
}

extern "C" int uwr_mdta_gram(const float* X, long long ldx, const float* Y, long long ldy, int B, int L, int heads,
                             int c, float* G, float* sqx, float* sqy, float* workspace, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_common("uwr_mdta_gram", B, L, heads, c)) return rc;
    UWR_REQUIRE(X && Y && G && workspace, "uwr_mdta_gram: null pointer");
    UWR_REQUIRE(heads * c * c <= 4096, "uwr_mdta_gram: heads * c * c must be <= 4096");
    UWR_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && (((uintptr_t)X | (uintptr_t)Y) & 15) == 0,
                "uwr_mdta_gram: operands must be 16-byte aligned with row strides in multiples of 4");
    switch (c) {
        case 8: return launch_gram<8>(X, ldx, Y, ldy, B, L, heads, G, sqx, sqy, workspace, stream);
        case 16: return launch_gram<16>(X, ldx, Y, ldy, B, L, heads, G, sqx, sqy, workspace, stream);
        case 32: return launch_gram<32>(X, ldx, Y, ldy, B, L, heads, G, sqx, sqy, workspace, stream);
        default: return launch_gram<64>(X, ldx, Y, ldy, B, L, heads, G, sqx, sqy, workspace, stream);
    }
}

extern "C" int uwr_mdta_apply(const float* X, long long ldx, const float* Mx, int transpose, const float* Yd,
                              long long ldy, const float* diag, float* out, long long ldo, int B, int L, int heads,
                              int c, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_common("uwr_mdta_apply", B, L, heads, c)) return rc;
    UWR_REQUIRE(X && Mx && out, "uwr_mdta_apply: null pointer");
    UWR_REQUIRE((diag == nullptr) == (Yd == nullptr), "uwr_mdta_apply: diag and Yd go together");
    UWR_REQUIRE(ldx % 4 == 0 && ldo % 2 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)out & 7) == 0 &&
                    (Yd == nullptr || (ldy % 2 == 0 && ((uintptr_t)Yd & 7) == 0)),
                "uwr_mdta_apply: alignment");
    switch (c) {
        case 8: return launch_apply<8>(X, ldx, Mx, transpose, Yd, ldy, diag, out, ldo, B, L, heads, stream);
        case 16: return launch_apply<16>(X, ldx, Mx, transpose, Yd, ldy, diag, out, ldo, B, L, heads, stream);
        case 32: return launch_apply<32>(X, ldx, Mx, transpose, Yd, ldy, diag, out, ldo, B, L, heads, stream);
        default: return launch_apply<64>(X, ldx, Mx, transpose, Yd, ldy, diag, out, ldo, B, L, heads, stream);
    }
}
