// TF32 tensor-core GEMM with fused epilogues for the Linear layers of the restoration
// transformers (AST.py:47-48,104,297,302,332,337).  Warp-level mma.sync.m16n8k8 pipeline fed by a
// 3-stage cp.async ring; see DESIGN.md §kernels for the tile/roofline discussion.
//
//   C[M,N] = epi( opA(A) * opB(B) )     128 x BN x 32 CTA tile, 8 warps (4 along M, 2 along N)
//
// Layouts: A_KM => A stored [K][M] (weight-gradient GEMMs, contraction over tokens),
//          B_NK => B stored [N][K] (nn.Linear weight layout).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int NTHREADS = 256;

struct GemmParams {
    const float* A;
    long long lda;
    const float* B;
    long long ldb;
    const float* B2;
    int n_split;
    float* C;
    long long ldc;
    int M, N, K;
    int k_per_split;  // contraction range handled by one blockIdx.z
    const float* bias;
    const float* bias2;
    const float* R;
    long long ldr;
    const float* rowscale;
    int rows_per_group;
    float* colsum;  // TN only: per-split partial column sums live at colsum + z*M
    long long c_split_stride;
    int round_out;
};

template <int BN, bool A_KM, bool B_NK>
struct TileCfg {
    static constexpr int A_STRIDE = A_KM ? (BM + 8) : (BK + 4);
    static constexpr int A_TILE = A_KM ? BK * A_STRIDE : BM * A_STRIDE;
    static constexpr int B_STRIDE = B_NK ? (BK + 4) : (BN + 8);
    static constexpr int B_TILE = B_NK ? BN * B_STRIDE : BK * B_STRIDE;
    static constexpr int SMEM_BYTES = STAGES * (A_TILE + B_TILE) * (int)sizeof(float);
};

template <int BN, bool A_KM, bool B_NK, int EPI, bool KSCALE, bool X3>
__global__ void __launch_bounds__(NTHREADS, X3 ? 1 : 2) gemm_tf32_kernel(const GemmParams p) {
    uwr_pdl_enter();
    using Cfg = TileCfg<BN, A_KM, B_NK>;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;
    float* Bs = smem + STAGES * Cfg::A_TILE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g = lane >> 2;  // group id 0..7
    const int t = lane & 3;   // thread in group 0..3
    const int wm = warp & 3;
    const int wn = warp >> 2;
    constexpr int WN_TILE = BN / 2;   // columns per warp
    constexpr int NT = WN_TILE / 8;   // n8 tiles per warp
    constexpr int MT = 2;             // m16 tiles per warp

    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int kbegin = blockIdx.z * p.k_per_split;
    const int kend = min(p.K, kbegin + p.k_per_split);
    const int KT = (kend - kbegin + BK - 1) / BK;

    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
    float csum[MT][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const bool do_colsum = A_KM && (p.colsum != nullptr) && (blockIdx.x == 0) && (wn == 0);

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kbegin + kt * BK;
        float* as = As + stage * Cfg::A_TILE;
        float* bs = Bs + stage * Cfg::B_TILE;
        // ---- A ----
#pragma unroll
        for (int i = 0; i < (BM * BK / 4) / NTHREADS; ++i) {
            const int chunk = tid + i * NTHREADS;
            if (!A_KM) {
                const int row = chunk >> 3, c4 = (chunk & 7) << 2;
                const int gm = m0 + row, gk = k0 + c4;
                const bool ok = (gm < p.M) && (gk < kend);
                const float* src = ok ? (p.A + (long long)gm * p.lda + gk) : p.A;
                cp_async16(as + row * Cfg::A_STRIDE + c4, src, ok);
            } else {
                const int krow = chunk >> 5, c4 = (chunk & 31) << 2;
                const int gk = k0 + krow, gm = m0 + c4;
                const bool ok = (gk < kend) && (gm < p.M);
                const float* src = ok ? (p.A + (long long)gk * p.lda + gm) : p.A;
                cp_async16(as + krow * Cfg::A_STRIDE + c4, src, ok);
            }
        }
        // ---- B ----
#pragma unroll
        for (int i = 0; i < (BN * BK / 4) / NTHREADS; ++i) {
            const int chunk = tid + i * NTHREADS;
            if (B_NK) {
                const int row = chunk >> 3, c4 = (chunk & 7) << 2;
                const int gn = n0 + row, gk = k0 + c4;
                const bool ok = (gn < p.N) && (gk < kend);
                const float* base = (gn < p.n_split) ? (p.B + (long long)gn * p.ldb)
                                                     : (p.B2 + (long long)(gn - p.n_split) * p.ldb);
                const float* src = ok ? (base + gk) : p.B;
                cp_async16(bs + row * Cfg::B_STRIDE + c4, src, ok);
            } else {
                constexpr int CPR = BN / 4;  // chunks per k-row
                const int krow = chunk / CPR, c4 = (chunk % CPR) << 2;
                const int gk = k0 + krow, gn = n0 + c4;
                const bool ok = (gk < kend) && (gn < p.N);
                const float* base = (gk < p.n_split) ? (p.B + (long long)gk * p.ldb)
                                                     : (p.B2 + (long long)(gk - p.n_split) * p.ldb);
                const float* src = ok ? (base + gn) : p.B;
                cp_async16(bs + krow * Cfg::B_STRIDE + c4, src, ok);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_tile(nk % STAGES, nk);
            cp_async_commit();
        }
        const float* as = As + (kt % STAGES) * Cfg::A_TILE;
        const float* bs = Bs + (kt % STAGES) * Cfg::B_TILE;
        const int kglob = kbegin + kt * BK;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 8) {
            uint32_t af[MT][4];
            uint32_t al[X3 ? MT : 1][4];  // low parts (3xTF32 mode)
            float ks0 = 1.f, ks1 = 1.f;
            if (KSCALE) {
                // DropPath scale of the incoming gradient rows (contraction index)
                const int ka = min(kglob + kk + t, p.K - 1), kb = min(kglob + kk + t + 4, p.K - 1);
                ks0 = __ldg(p.rowscale + ka / p.rows_per_group);
                ks1 = __ldg(p.rowscale + kb / p.rows_per_group);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r0 = wm * 32 + mt * 16;
                float a0, a1, a2, a3;
                if (!A_KM) {
                    a0 = as[(r0 + g) * Cfg::A_STRIDE + kk + t];
                    a1 = as[(r0 + g + 8) * Cfg::A_STRIDE + kk + t];
                    a2 = as[(r0 + g) * Cfg::A_STRIDE + kk + t + 4];
                    a3 = as[(r0 + g + 8) * Cfg::A_STRIDE + kk + t + 4];
                } else {
                    a0 = as[(kk + t) * Cfg::A_STRIDE + r0 + g];
                    a1 = as[(kk + t) * Cfg::A_STRIDE + r0 + g + 8];
                    a2 = as[(kk + t + 4) * Cfg::A_STRIDE + r0 + g];
                    a3 = as[(kk + t + 4) * Cfg::A_STRIDE + r0 + g + 8];
                }
                if (KSCALE) {
                    a0 *= ks0;
                    a1 *= ks0;
                    a2 *= ks1;
                    a3 *= ks1;
                }
                if (A_KM) {
                    if (do_colsum) {
                        csum[mt][0] += a0 + a2;
                        csum[mt][1] += a1 + a3;
                    }
                }
                af[mt][0] = f2tf32(a0);
                af[mt][1] = f2tf32(a1);
                af[mt][2] = f2tf32(a2);
                af[mt][3] = f2tf32(a3);
                if (X3) {
                    al[X3 ? mt : 0][0] = f2tf32(a0 - __uint_as_float(af[mt][0]));
                    al[X3 ? mt : 0][1] = f2tf32(a1 - __uint_as_float(af[mt][1]));
                    al[X3 ? mt : 0][2] = f2tf32(a2 - __uint_as_float(af[mt][2]));
                    al[X3 ? mt : 0][3] = f2tf32(a3 - __uint_as_float(af[mt][3]));
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int c0 = wn * WN_TILE + nt * 8;
                uint32_t bf[2];
                float b0, b1;
                if (B_NK) {
                    b0 = bs[(c0 + g) * Cfg::B_STRIDE + kk + t];
                    b1 = bs[(c0 + g) * Cfg::B_STRIDE + kk + t + 4];
                } else {
                    b0 = bs[(kk + t) * Cfg::B_STRIDE + c0 + g];
                    b1 = bs[(kk + t + 4) * Cfg::B_STRIDE + c0 + g];
                }
                bf[0] = f2tf32(b0);
                bf[1] = f2tf32(b1);
                if (X3) {
                    uint32_t bl[2];
                    bl[0] = f2tf32(b0 - __uint_as_float(bf[0]));
                    bl[1] = f2tf32(b1 - __uint_as_float(bf[1]));
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_tf32_16x8x8(acc[mt][nt], al[X3 ? mt : 0], bf);
                        mma_tf32_16x8x8(acc[mt][nt], af[mt], bl);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_tf32_16x8x8(acc[mt][nt], af[mt], bf);
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----
    float* Cout = p.C + (long long)blockIdx.z * p.c_split_stride;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = m0 + wm * 32 + mt * 16 + g + half * 8;
            if (row >= p.M) continue;
            float s = 1.f;
            if (!A_KM && p.rowscale != nullptr) s = __ldg(p.rowscale + row / p.rows_per_group);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = n0 + wn * WN_TILE + nt * 8 + 2 * t;
                if (col >= p.N) continue;
                float v0 = acc[mt][nt][half * 2 + 0];
                float v1 = acc[mt][nt][half * 2 + 1];
                if (p.bias != nullptr) {
                    const float* bp = (col < p.n_split) ? (p.bias + col) : (p.bias2 + (col - p.n_split));
                    v0 += __ldg(bp);
                    v1 += __ldg(bp + 1);
                }
                v0 *= s;
                v1 *= s;
                if (EPI == UWR_EPI_RESID) {
                    const float2 r = *reinterpret_cast<const float2*>(p.R + (long long)row * p.ldr + col);
                    v0 += r.x;
                    v1 += r.y;
                } else if (EPI == UWR_EPI_MUL_DGELU) {
                    const float2 r = *reinterpret_cast<const float2*>(p.R + (long long)row * p.ldr + col);
                    v0 *= gelu_grad_f(r.x);
                    v1 *= gelu_grad_f(r.y);
                } else if (EPI == UWR_EPI_MUL) {
                    const float2 r = *reinterpret_cast<const float2*>(p.R + (long long)row * p.ldr + col);
                    v0 *= r.x;
                    v1 *= r.y;
                }
                if (p.round_out) {
                    v0 = tf32_round(v0);
                    v1 = tf32_round(v1);
                }
                *reinterpret_cast<float2*>(Cout + (long long)row * p.ldc + col) = make_float2(v0, v1);
            }
        }
    }
    if (A_KM) {
        if (do_colsum) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v = csum[mt][half];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int row = m0 + wm * 32 + mt * 16 + g + half * 8;
                    if (t == 0 && row < p.M) p.colsum[(long long)blockIdx.z * p.M + row] = v;
                }
        }
    }
}

// sum split-K partials: out[i] = sum_z ws[z*stride + i]
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out,
                                     long long n, long long stride, int splits, long long ld_out,
                                     int ncols) {
    uwr_pdl_enter();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += ws[z * stride + i];
        const long long r = i / ncols, c = i % ncols;
        out[r * ld_out + c] = s;
    }
}

struct SplitPlan {
    int splits;
    int k_per_split;
};

SplitPlan plan_split(int M, int N, int K, int a_km, int bn) {
    SplitPlan sp{1, K};
    if (!a_km) return sp;
    const long long tiles = (long long)uwr_cdiv(M, BM) * uwr_cdiv(N, bn);
    const int target = 2 * uwr_sm_count();
    int splits = (int)((target + tiles - 1) / tiles);
    const int max_splits = K / 256 > 0 ? K / 256 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 512) splits = 512;
    int kper = uwr_cdiv(K, splits);
    kper = ((kper + BK - 1) / BK) * BK;
    splits = uwr_cdiv(K, kper);
    sp.splits = splits;
    sp.k_per_split = kper;
    return sp;
}

int pick_bn(int N) {
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    return 128;
}


template <int BN, bool A_KM, bool B_NK, int EPI, bool KSCALE, bool X3>
int launch_x(const GemmParams& p, int splits, cudaStream_t stream) {
    using Cfg = TileCfg<BN, A_KM, B_NK>;
    auto kern = gemm_tf32_kernel<BN, A_KM, B_NK, EPI, KSCALE, X3>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid(uwr_cdiv(p.N, BN), uwr_cdiv(p.M, BM), splits);
    (void)uwr_launch_pdl(kern, dim3(grid), dim3(NTHREADS), Cfg::SMEM_BYTES, stream, p);
    UWR_CHECK_LAUNCH("gemm_tf32_kernel");
    return 0;
}

template <int BN, bool A_KM, bool B_NK, int EPI, bool KSCALE>
int launch(const GemmParams& p, int splits, cudaStream_t stream) {
    return g_uwr_gemm_passes == 3 ? launch_x<BN, A_KM, B_NK, EPI, KSCALE, true>(p, splits, stream)
                         : launch_x<BN, A_KM, B_NK, EPI, KSCALE, false>(p, splits, stream);
}

template <int BN>
int dispatch_layout(const uwr_gemm_desc* d, const GemmParams& p, int splits, cudaStream_t stream) {
    const bool kscale = d->a_km && d->rowscale != nullptr;
    if (d->a_km) {
        UWR_REQUIRE(!d->b_nk, "uwr_gemm_tf32: a_km=1 requires b_nk=0");
        UWR_REQUIRE(d->epilogue == UWR_EPI_NONE, "uwr_gemm_tf32: a_km=1 supports no epilogue");
        return kscale ? launch<BN, true, false, UWR_EPI_NONE, true>(p, splits, stream)
                      : launch<BN, true, false, UWR_EPI_NONE, false>(p, splits, stream);
    }
    if (d->b_nk) {
        switch (d->epilogue) {
            case UWR_EPI_NONE: return launch<BN, false, true, UWR_EPI_NONE, false>(p, splits, stream);
            case UWR_EPI_RESID: return launch<BN, false, true, UWR_EPI_RESID, false>(p, splits, stream);
            case UWR_EPI_MUL_DGELU: return launch<BN, false, true, UWR_EPI_MUL_DGELU, false>(p, splits, stream);
            case UWR_EPI_MUL: return launch<BN, false, true, UWR_EPI_MUL, false>(p, splits, stream);
        }
    } else {
        switch (d->epilogue) {
            case UWR_EPI_NONE: return launch<BN, false, false, UWR_EPI_NONE, false>(p, splits, stream);
            case UWR_EPI_RESID: return launch<BN, false, false, UWR_EPI_RESID, false>(p, splits, stream);
            case UWR_EPI_MUL_DGELU: return launch<BN, false, false, UWR_EPI_MUL_DGELU, false>(p, splits, stream);
            case UWR_EPI_MUL: return launch<BN, false, false, UWR_EPI_MUL, false>(p, splits, stream);
        }
    }
    uwr_set_error("uwr_gemm_tf32: bad epilogue %d", d->epilogue);
    return -1;
}

}  // namespace

extern "C" size_t uwr_gemm_workspace_bytes(int M, int N, int K, int a_km) {
    if (!a_km) return 0;
    const SplitPlan sp = plan_split(M, N, K, a_km, pick_bn(N));
    if (sp.splits <= 1) return 0;
    return (size_t)sp.splits * ((size_t)M * N + (size_t)M) * sizeof(float);
}

extern "C" int uwr_gemm_tf32(const uwr_gemm_desc* d, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(d && d->A && d->B && d->C, "uwr_gemm_tf32: null operand");
    UWR_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "uwr_gemm_tf32: empty problem %d %d %d", d->M, d->N, d->K);
    UWR_REQUIRE(!d->c_half && !d->r_half, "uwr_gemm_tf32: fp16 C / R storage is served by uwr_gemm_tcgen05 only");
    UWR_REQUIRE(d->K % 4 == 0 && d->lda % 4 == 0 && d->ldb % 4 == 0 && d->ldc % 2 == 0 && d->N % 2 == 0,
                "uwr_gemm_tf32: K,lda,ldb must be multiples of 4 and N,ldc even (K=%d lda=%lld ldb=%lld ldc=%lld N=%d)",
                d->K, d->lda, d->ldb, d->ldc, d->N);
    UWR_REQUIRE(((uintptr_t)d->A % 16 == 0) && ((uintptr_t)d->B % 16 == 0) && ((uintptr_t)d->C % 8 == 0),
                "uwr_gemm_tf32: operands must be 16-byte aligned");
    if (d->a_km) UWR_REQUIRE(d->M % 4 == 0, "uwr_gemm_tf32: a_km=1 needs M %% 4 == 0");
    if (!d->b_nk) UWR_REQUIRE(d->N % 4 == 0, "uwr_gemm_tf32: b_nk=0 needs N %% 4 == 0");
    if (d->B2) UWR_REQUIRE(d->n_split > 0 && d->n_split % 4 == 0 && ((uintptr_t)d->B2 % 16 == 0),
                           "uwr_gemm_tf32: segmented B needs n_split %% 4 == 0 and 16-byte alignment");
    if (d->B2 && d->bias) UWR_REQUIRE(d->b_nk, "uwr_gemm_tf32: segmented bias needs b_nk=1");
    if (d->epilogue != UWR_EPI_NONE) UWR_REQUIRE(d->R && d->ldr % 2 == 0, "uwr_gemm_tf32: epilogue needs R");
    if (d->rowscale) UWR_REQUIRE(d->rows_per_group > 0, "uwr_gemm_tf32: rowscale needs rows_per_group");

    const int bn = pick_bn(d->N);
    const SplitPlan sp = plan_split(d->M, d->N, d->K, d->a_km, bn);

    GemmParams p;
    p.A = d->A; p.lda = d->lda; p.B = d->B; p.ldb = d->ldb;
    p.B2 = d->B2 ? d->B2 : d->B;
    p.n_split = d->B2 ? d->n_split : 0x7fffffff;
    p.C = d->C; p.ldc = d->ldc; p.M = d->M; p.N = d->N; p.K = d->K;
    p.k_per_split = sp.k_per_split;
    p.bias = d->bias; p.bias2 = d->bias2 ? d->bias2 : d->bias;
    p.R = d->R; p.ldr = d->ldr;
    p.rowscale = d->rowscale; p.rows_per_group = d->rows_per_group > 0 ? d->rows_per_group : 1;
    p.colsum = d->colsum; p.c_split_stride = 0;
    p.round_out = d->round_out;

    float* ws_c = nullptr;
    float* ws_cs = nullptr;
    if (sp.splits > 1) {
        const size_t need = (size_t)sp.splits * ((size_t)d->M * d->N + (size_t)d->M) * sizeof(float);
        UWR_REQUIRE(d->workspace && d->workspace_bytes >= need, "uwr_gemm_tf32: workspace too small (%zu < %zu)",
                    d->workspace_bytes, need);
        ws_c = d->workspace;
        ws_cs = d->workspace + (size_t)sp.splits * d->M * d->N;
        p.C = ws_c; p.ldc = d->N; p.c_split_stride = (long long)d->M * d->N;
        if (d->colsum) p.colsum = ws_cs;
    }

    int rc;
    switch (bn) {
        case 32: rc = dispatch_layout<32>(d, p, sp.splits, stream); break;
        case 64: rc = dispatch_layout<64>(d, p, sp.splits, stream); break;
        default: rc = dispatch_layout<128>(d, p, sp.splits, stream); break;
    }
    if (rc) return rc;

    if (sp.splits > 1) {
        const long long n = (long long)d->M * d->N;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 4 * uwr_sm_count()) blocks = 4 * uwr_sm_count();
        (void)uwr_launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, ws_c, d->C, n, n, sp.splits, d->ldc, d->N);
        UWR_CHECK_LAUNCH("splitk_reduce_kernel");
        if (d->colsum) {
            (void)uwr_launch_pdl(splitk_reduce_kernel, dim3(uwr_cdiv(d->M, 256)), dim3(256), 0, stream, ws_cs, d->colsum, d->M, d->M, sp.splits,
                                                                         d->M, d->M);
            UWR_CHECK_LAUNCH("splitk_reduce_kernel(colsum)");
        }
    }
    return 0;
}
