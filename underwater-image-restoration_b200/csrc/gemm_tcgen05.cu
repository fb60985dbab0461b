// Blackwell-native streaming GEMM for the Linear layers (AST.py:47-48,104,297,302,332,337):
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma kind::tf32 issued by
// one thread, fp32 accumulators in TMEM (two buffers) -> tcgen05.ld epilogue warps -> global.
//
//   persistent CTAs (one per SM), 320 threads:
//     warp 0   : TMA producer          (one elected lane)
//     warp 1   : TMEM allocator + MMA issuer (one elected lane)
//     warp 2-9 : epilogue, warp w owns TMEM lanes 32*(w%4) .. +31 of the 128-row accumulator and
//                every second 32-column chunk; bias staged in smem, R prefetched, TMEM loads pipelined
//
//   CTA tile 128 x BN (BN <= 256), contraction streamed in chunks of 32 fp32 (128 B rows).
//   Layouts (same meaning as uwr_gemm_tf32):  NT: A[M][K], B[N][K]   (forward,   both K-major)
//                                              NN: A[M][K], B[K][N]   (data grad, B MN-major)
//                                              TN: A[K][M], B[K][N]   (weight grad, both MN-major,
//                                                                      split over the contraction)
// Operand contract: kind::tf32 reads the upper 19 bits of each fp32 (truncation), so callers pass
// operands already rounded to TF32 (the producing kernels round at their stores, weights through
// uwr_round_tf32_tensors); the GEMM itself adds no rounding bias.
#include <cuda.h>
#include <cuda_fp16.h>

#include "uwr_common.cuh"
#include "uwr_tma.cuh"
#include "../../include/uwr_b200.h"

namespace {

using namespace uwr_tma;

constexpr int TM = 128;      // CTA tile rows (UMMA M)
constexpr int KC = 32;       // contraction chunk: 32 fp32 = one 128-byte swizzle row
constexpr int EP_STRIDE = 36;  // fp32 words per staged epilogue row (32 + 4: conflict-free float4 access)
constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quadrant, interleaved over 32-column chunks
constexpr int T5_THREADS = 64 + EPI_WARPS * 32;

enum { LAY_NT = 0, LAY_NN = 1, LAY_TN = 2 };

// Virtual im2col operand (template parameter CV): the token matrix x (B*H*W, ld) is addressed as a 4-D (C, W, H, B)
// tensor and each 32-channel K chunk (CV = 1: A operand, rows = 128 output pixels) or 32-wide column group (CV = 2: B
// operand of the weight gradient, rows = 32 output pixels) is ONE box at the tap's offset -- out-of-image pixels are
// zero-filled by TMA (the padding), a stride-2 convolution uses the map's element strides.  No im2col buffer.
struct T5Conv {
    int cpt;      // 32-channel chunks per tap (Cin / 32)
    int kw;       // kernel width: tap -> (ky, kx) = (tap / kw, tap % kw)
    int taps;     // kh * kw
    int pad, stride;
    int ow, ohw;  // output width, output pixels per image
    int pw, ph;   // CV = 2: the 32 output pixels of one K chunk are a pw x ph patch
    int cin;
    int nb;       // batch: a coordinate >= nb in the outermost dimension reads zeros (padding column groups)
};

struct T5Params {
    T5Conv cv;
    float* C;
    long long ldc;
    int M, N;            // output extent
    int chunks;          // contraction chunks (of 32) in total
    int chunks_per_split;
    int splits;
    long long split_stride;  // elements between split partials (TN)
    int tiles_m, tiles_n;
    const float* bias;
    const float* R;
    long long ldr;
    const float* rowscale;
    int rows_per_group;
    int epilogue;
    int round_out;
};

// two floats -> packed __half2 bits, clamped to the finite fp16 range (|x| > 65504 would convert to inf)
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------- PTX wrappers
// -------------------------------------------------------------------------------------- kernel
template <int BN>
// Ring depth: what matters is the BYTES in flight per SM (HBM latency x the SM's share of the bandwidth is ~45 KB).  A stage
// of a narrow tile carries little -- BN = 32: 16 KB of A + 4 KB of B, and for a thin weight gradient (M = 32) only 4 KB of
// that A tile is inside the matrix -- so narrow tiles get a deeper ring (same box: SpectralTransformer, whose GEMMs are
// mostly N, M <= 64, 228 -> 238 img/s; AST +0.4 %; one stage more again changed nothing).
__host__ __device__ constexpr int t5_stages() { return BN >= 256 ? 3 : (BN >= 128 ? 4 : (BN >= 64 ? 6 : 8)); }
template <int BN>
__host__ __device__ constexpr int t5_smem_bytes() {
    return t5_stages<BN>() * (TM * KC * 4 + BN * KC * 4) + 256 + BN * 4 + EPI_WARPS * 32 * EP_STRIDE * 4 + 1024;
}

// HF: fp16 storage flags, bit 0 = C is __half (EPI_NONE only), bit 1 = R is __half.
// CL: thread-block cluster size along M (1 or 2).  With CL = 2 the two CTAs of a cluster work on neighbouring row
// tiles of the SAME column tile: each TMA-loads half of the B tile and MULTICASTS it into both CTAs' shared memory, so
// the L2 -> SM operand traffic per CTA and contraction chunk drops from A + B to A + B/2 (48 -> 32 KB at BN = 256) —
// the tensor-bound shapes are limited by exactly that traffic.  A stage may only be refilled when BOTH CTAs' MMAs have
// consumed it: every CTA's tcgen05.commit arrives on the stage's empty barrier in both CTAs (count CL).
template <int BN, int LAY, int EPI, int HF = 0, int CL = 1, int CV = 0>
__global__ void __launch_bounds__(T5_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const T5Params p) {
    constexpr int STAGES = t5_stages<BN>();
    constexpr bool A_MN = (LAY == LAY_TN);
    constexpr bool B_MN = (LAY != LAY_NT);
    constexpr bool C_HALF = (HF & 1) != 0, R_HALF = (HF & 2) != 0;
    constexpr int A_BYTES = TM * KC * 4;  // 16 KB
    constexpr int B_BYTES = BN * KC * 4;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two: BN in {32,64,128,256}
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) |
                               ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // 1024 B: swizzle atoms
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sbias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [BN]
    float* sstage = sbias + BN;  // [EPI_WARPS][32][EP_STRIDE] accumulator transposition buffers

    uwr_pdl_trigger();   // PDL: the next grid may be launched; the set-up below overlaps the predecessor's tail
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work items are enumerated per CLUSTER: (row-tile group, column tile, split); CTA `crank` takes row tile CL*g + crank
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int tiles_mg = (p.tiles_m + CL - 1) / CL;
    const int items = tiles_mg * p.tiles_n * p.splits;
    const int first_item = (int)blockIdx.x / CL, item_step = (int)gridDim.x / CL;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CL);   // one tcgen05.commit arrival per CTA of the cluster
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // barriers of every CTA are initialised before any remote TMA / commit touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uwr_pdl_wait();      // the predecessor grid has completed: global memory may be read and written from here on

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int item = first_item; item < items; item += item_step) {
                const int tm = (item % tiles_mg) * CL + crank;   // may be >= tiles_m for the last group: TMA zero-fills
                const int tn = (item / tiles_mg) % p.tiles_n;
                const int z = item / (tiles_mg * p.tiles_n);
                const int c_begin = z * p.chunks_per_split;
                const int c_end = min(p.chunks, c_begin + p.chunks_per_split);
                // virtual im2col operands: all index divisions happen once per work item, the chunk loop only steps
                // (the single producer thread must stay far below one MMA chunk, 512 cycles, per iteration)
                int cvb = 0, cvx = 0, cvy = 0;          // CV 1: image / origin of the tile; CV 2: of the current chunk
                int cvc0 = 0, cvkx = 0, cvky = 0;       // CV 1: channel offset and tap of the current chunk
                constexpr int NG = CV == 2 ? BN / 32 : (CV == 3 ? TM / 32 : 1);   // 32-wide groups of the virtual operand
                int gx[NG], gy[NG], gc[NG];
                if (CV == 1) {
                    const int m0 = tm * TM, r = m0 % p.cv.ohw, oy0 = r / p.cv.ow;
                    cvb = m0 / p.cv.ohw;
                    cvx = (r - oy0 * p.cv.ow) * p.cv.stride - p.cv.pad;
                    cvy = oy0 * p.cv.stride - p.cv.pad;
                    const int tap = c_begin / p.cv.cpt;
                    cvc0 = (c_begin - tap * p.cv.cpt) * KC;
                    cvky = tap / p.cv.kw;
                    cvkx = tap - cvky * p.cv.kw;
                }
                if (CV == 2 || CV == 3) {
                    const int k0 = c_begin * KC, r = k0 % p.cv.ohw;
                    cvb = k0 / p.cv.ohw;
                    cvy = r / p.cv.ow;                  // in output pixels; scaled where the box is issued
                    cvx = r - cvy * p.cv.ow;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const int n0 = (CV == 2 ? tn * BN : tm * TM) + g * 32, tap = n0 / p.cv.cin, ky = tap / p.cv.kw;
                        gc[g] = tap < p.cv.taps ? n0 - tap * p.cv.cin : -1;   // -1: padding group beyond the last tap
                        gy[g] = ky - p.cv.pad;
                        gx[g] = tap - ky * p.cv.kw - p.cv.pad;
                    }
                }
                for (int c = c_begin; c < c_end; ++c) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    if (CV == 1) {
                        // virtual im2col rows: the tile's 128 output pixels (one box) at tap (ky, kx), channels c0..c0+31
                        tma_load_4d(sa, &mapA, &full_bar[stage], cvc0, cvx + cvkx, cvy + cvky, cvb);
                        cvc0 += KC;
                        if (cvc0 == p.cv.cin) {
                            cvc0 = 0;
                            if (++cvkx == p.cv.kw) {
                                cvkx = 0;
                                ++cvky;
                            }
                        }
                    } else if (CV == 3) {
                        // virtual im2col as the MN-major A operand (transposed weight gradient): 32-pixel patch x 32-channel group
                        const int xs = cvx * p.cv.stride, ys = cvy * p.cv.stride;
#pragma unroll
                        for (int g = 0; g < TM / 32; ++g)
                            tma_load_4d(sa + g * 4096, &mapA, &full_bar[stage], gc[g] < 0 ? 0 : gc[g], xs + gx[g], ys + gy[g],
                                        gc[g] < 0 ? p.cv.nb : cvb);
                    } else if (!A_MN) {
                        tma_load_2d(sa, &mapA, &full_bar[stage], c * KC, tm * TM);
                    } else {
                        // MN-major: one 32(mn) x 32(k) box per 32-wide group -> [group][k][128 B]
#pragma unroll
                        for (int g = 0; g < TM / 32; ++g)
                            tma_load_2d(sa + g * 4096, &mapA, &full_bar[stage], tm * TM + g * 32, c * KC);
                    }
                    if (CV == 2) {
                        // virtual im2col columns: this chunk's 32 output pixels x the 32-channel group of tap (ky, kx)
                        const int xs = cvx * p.cv.stride, ys = cvy * p.cv.stride;
#pragma unroll
                        for (int g = 0; g < BN / 32; ++g)
                            tma_load_4d(sb + g * 4096, &mapB, &full_bar[stage], gc[g] < 0 ? 0 : gc[g], xs + gx[g], ys + gy[g],
                                        gc[g] < 0 ? p.cv.nb : cvb);
                    } else if (CL == 1) {
                        if (!B_MN) {
                            tma_load_2d(sb, &mapB, &full_bar[stage], c * KC, tn * BN);
                        } else {
#pragma unroll
                            for (int g = 0; g < BN / 32; ++g)
                                tma_load_2d(sb + g * 4096, &mapB, &full_bar[stage], tn * BN + g * 32, c * KC);
                        }
                    } else {
                        // this CTA's share of the B tile (mapB's box is BN / CL rows here), multicast to the whole cluster
                        constexpr int SHARE = BN / CL;
                        if (!B_MN) {
                            tma_load_2d_mc(sb + crank * SHARE * 128, &mapB, &full_bar[stage], c * KC, tn * BN + crank * SHARE,
                                           CMASK);
                        } else {
#pragma unroll
                            for (int g = 0; g < SHARE / 32; ++g) {
                                const int gg = crank * (SHARE / 32) + g;
                                tma_load_2d_mc(sb + gg * 4096, &mapB, &full_bar[stage], tn * BN + gg * 32, c * KC, CMASK);
                            }
                        }
                    }
                    if (CV == 2 || CV == 3) {
                        cvx += p.cv.pw;                  // next 32-pixel patch: along the row, then down, then the next image
                        if (cvx == p.cv.ow) {
                            cvx = 0;
                            cvy += p.cv.ph;
                            if (cvy * p.cv.ow == p.cv.ohw) {
                                cvy = 0;
                                ++cvb;
                            }
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int item = first_item; item < items; item += item_step, ++local) {
                const int z = item / (tiles_mg * p.tiles_n);
                const int c_begin = z * p.chunks_per_split;
                const int c_end = min(p.chunks, c_begin + p.chunks_per_split);
                const int buf = local & 1;
                const uint32_t use = (uint32_t)(local >> 1);  // how often this buffer was used before
                mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int c = c_begin; c < c_end; ++c) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k8 = 0; k8 < KC / 8; ++k8) {
                        // K-major: rows of 128 B, 8-row groups 1024 B apart, step 32 B inside the row.
                        // MN-major: [mn/32][k][32 floats]; 4-k atoms 512 B apart (SBO), mn groups 4096 B
                        // apart (LBO); one MMA consumes 8 k rows = 1024 B.
                        const uint64_t ad = A_MN ? make_smem_desc(sa + k8 * 1024, 4096, 512, 1)
                                                 : make_smem_desc(sa + k8 * 32, 16, 1024, 2);
                        const uint64_t bd = B_MN ? make_smem_desc(sb + k8 * 1024, 4096, 512, 1)
                                                 : make_smem_desc(sb + k8 * 32, 16, 1024, 2);
                        umma_tf32(d_tmem, ad, bd, IDESC, (c > c_begin || k8 > 0) ? 1u : 0u);
                    }
                    // frees the smem slot once these MMAs retire — in every CTA that multicasts into it
                    if (CL == 1) umma_commit(&empty_bar[stage]);
                    else umma_commit_mc(&empty_bar[stage], CMASK);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[buf]);  // accumulator complete
            }
        }
    } else {
        // ================================ epilogue ====================================
        const int q = warp & 3;              // TMEM lane quadrant owned by this warp
        const int half = (warp - 2) >> 2;    // 0/1: which 32-column chunks (even/odd)
        const int etid = threadIdx.x - 64;   // 0..255 inside the epilogue group
        int local = 0, cur_tn = -1;
        for (int item = first_item; item < items; item += item_step, ++local) {
            const int tm = (item % tiles_mg) * CL + crank;
            const int tn = (item / tiles_mg) % p.tiles_n;
            const int z = item / (tiles_mg * p.tiles_n);
            const int buf = local & 1;
            const uint32_t use = (uint32_t)(local >> 1);
            if (p.bias != nullptr && tn != cur_tn) {  // uniform across the epilogue warps
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                for (int i = etid; i < BN; i += EPI_WARPS * 32) {
                    const int col = tn * BN + i;
                    sbias[i] = col < p.N ? __ldg(p.bias + col) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            }
            cur_tn = tn;
            // Accumulator rows arrive one per lane (TMEM lane = tile row).  They are transposed through
            // a per-warp smem buffer so that every global access of the epilogue (C stores, R loads)
            // covers 4 rows x 128 contiguous bytes per warp instruction instead of 32 rows x 16 bytes.
            float* st = sstage + (warp - 2) * (32 * EP_STRIDE);
            // fp32 C: lane -> (row within a group of 4, float4 of the 32-column chunk): 4 rows x 128 B per instruction.
            // fp16 C: lane -> (row within a group of 8, 8 columns): 8 rows x 64 B per instruction, 16 B per lane.
            const int rsub = C_HALF ? (lane >> 2) : (lane >> 3);
            const int cc = C_HALF ? (lane & 3) * 8 : (lane & 7) * 4;
            constexpr int RGROUPS = C_HALF ? 4 : 8;          // row groups per 32-row chunk
            constexpr int RSTEP = C_HALF ? 8 : 4;
            const int row0 = tm * TM + q * 32;
            float sc[RGROUPS];
#pragma unroll
            for (int i = 0; i < RGROUPS; ++i) {
                const int row = row0 + i * RSTEP + rsub;
                sc[i] = (p.rowscale != nullptr && row < p.M) ? __ldg(p.rowscale + row / p.rows_per_group) : 1.f;
            }
            float* cbase = p.C + (long long)z * p.split_stride;
            constexpr int NCH = (BN / 32 + 1) / 2;  // chunks per warp (BN = 32: only half 0 works)
            // fp16 epilogue operand: 8 B per lane and row, so the loads of ALL of this warp's chunks of the tile fit in
            // registers (NCH x 8 x 8 B) and are issued before the accumulator is even ready: they land under the MMAs
            uint2 rraw_all[R_HALF ? NCH : 1][8];
            if (R_HALF) {
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c0 = (2 * ci + half) * 32;
                    const int col = tn * BN + c0 + cc;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row0 + i * 4 + rsub;
                        const bool ok = c0 < BN && col < p.N && row < p.M;
                        rraw_all[ci][i] = ok ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.R) +
                                                                               (long long)row * p.ldr + col)
                                             : make_uint2(0u, 0u);
                    }
                }
            }
            mbar_wait(&tfull_bar[buf], use & 1);
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
            uint32_t acc[2][32];
            if (half * 32 < BN) tmem_ld32_issue(tbase + half * 32, acc[0]);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = (2 * ci + half) * 32;
                if (c0 >= BN) break;
                const int col = tn * BN + c0 + cc;
                const bool col_ok = col < p.N;  // N is a multiple of 4 (8 with fp16 C)
                // operand of the epilogue (residual / multiplier): issue the loads before waiting on TMEM; fp16
                // operands stay raw (8 B) until they are used, so that the 8 loads are in flight together
                float4 rv[8];
                if (EPI != UWR_EPI_NONE && !R_HALF) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row0 + i * 4 + rsub;
                        rv[i] = (col_ok && row < p.M) ? *reinterpret_cast<const float4*>(p.R + (long long)row * p.ldr + col)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), b4b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) {
                    b4 = *reinterpret_cast<const float4*>(sbias + c0 + cc);
                    if (C_HALF) b4b = *reinterpret_cast<const float4*>(sbias + c0 + cc + 4);
                }
                tmem_ld_wait();
                {
                    const uint32_t* a = acc[ci & 1];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(st + lane * EP_STRIDE + 4 * j) =
                            make_uint4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                }
                const int cn = c0 + 64;
                if (ci + 1 < NCH && cn < BN) tmem_ld32_issue(tbase + cn, acc[(ci + 1) & 1]);  // next chunk in flight
                __syncwarp();
                if (C_HALF) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + rsub;
                        const int row = row0 + rr;
                        float4 o0 = *reinterpret_cast<const float4*>(st + rr * EP_STRIDE + cc);
                        float4 o1 = *reinterpret_cast<const float4*>(st + rr * EP_STRIDE + cc + 4);
                        o0.x = (o0.x + b4.x) * sc[i]; o0.y = (o0.y + b4.y) * sc[i];
                        o0.z = (o0.z + b4.z) * sc[i]; o0.w = (o0.w + b4.w) * sc[i];
                        o1.x = (o1.x + b4b.x) * sc[i]; o1.y = (o1.y + b4b.y) * sc[i];
                        o1.z = (o1.z + b4b.z) * sc[i]; o1.w = (o1.w + b4b.w) * sc[i];
                        if (col_ok && row < p.M) {
                            __half* dst = reinterpret_cast<__half*>(cbase) + (long long)row * p.ldc + col;
                            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_half2(o0.x, o0.y), pack_half2(o0.z, o0.w),
                                                                        pack_half2(o1.x, o1.y), pack_half2(o1.z, o1.w));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + rsub;
                        const int row = row0 + rr;
                        float4 o = *reinterpret_cast<const float4*>(st + rr * EP_STRIDE + cc);
                        o.x = (o.x + b4.x) * sc[i]; o.y = (o.y + b4.y) * sc[i];
                        o.z = (o.z + b4.z) * sc[i]; o.w = (o.w + b4.w) * sc[i];
                        if (EPI != UWR_EPI_NONE && R_HALF) {
                            const uint2 raw = rraw_all[R_HALF ? ci : 0][i];
                            const float2 r01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                            const float2 r23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                            rv[i] = make_float4(r01.x, r01.y, r23.x, r23.y);
                        }
                        if (EPI == UWR_EPI_RESID) {
                            o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w;
                        } else if (EPI == UWR_EPI_MUL) {
                            o.x *= rv[i].x; o.y *= rv[i].y; o.z *= rv[i].z; o.w *= rv[i].w;
                        } else if (EPI == UWR_EPI_MUL_DGELU) {
                            o.x *= gelu_grad_f(rv[i].x); o.y *= gelu_grad_f(rv[i].y);
                            o.z *= gelu_grad_f(rv[i].z); o.w *= gelu_grad_f(rv[i].w);
                        }
                        if (p.round_out) o = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
                        if (col_ok && row < p.M) *reinterpret_cast<float4*>(cbase + (long long)row * p.ldc + col) = o;
                    }
                }
                __syncwarp();  // the staging buffer is reused by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer can still multicast into it or arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// Sum of the per-split partial tiles (deterministic: fixed assignment, fixed order).  A CTA = 32 float4 columns x 8 split
// groups: the partials of a thin weight gradient are only a few thousand float4 wide, so one thread per column (the first
// version) walked up to 148 splits as a chain of dependent-latency loads with ~4 in flight -- 14 us for 10 MB.  Eight
// threads per column, four loads in flight each, then a fixed-order fold through shared memory.
constexpr int RED_GROUPS = 8;
__global__ void __launch_bounds__(32 * RED_GROUPS) t5_splitk_reduce_kernel(const float* __restrict__ ws,
                                                                          float* __restrict__ out, long long n,
                                                                          long long stride, int splits) {
    __shared__ float4 sh[RED_GROUPS][32];
    uwr_pdl_enter();
    const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long long cols = n / 4;
    for (long long base = (long long)blockIdx.x * 32; base < cols; base += (long long)gridDim.x * 32) {
        const long long i = (base + col) * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (base + col < cols) {
#pragma unroll 4
            for (int z = grp; z < splits; z += RED_GROUPS) {
                const float4 v = *reinterpret_cast<const float4*>(ws + z * stride + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        sh[grp][col] = s;
        __syncthreads();
        if (grp == 0 && base + col < cols) {
#pragma unroll
            for (int g = 1; g < RED_GROUPS; ++g) {
                const float4 v = sh[g][col];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            *reinterpret_cast<float4*>(out + i) = s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ host side
int encode_2d(CUtensorMap* m, const float* base, long long inner, long long rows, long long ld, int box_rows,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {KC, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        uwr_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return -3;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        uwr_set_error("cuTensorMapEncodeTiled(2d) failed: %d (inner %lld rows %lld ld %lld)", (int)r, inner, rows, ld);
        return -3;
    }
    return 0;
}
// MN-major operand stored [k][width]: plain 2-D map, box = 32 (width) x 32 (k rows)
int encode_mn(CUtensorMap* m, const float* base, long long width, long long krows, long long ld) {
    return encode_2d(m, base, width, krows, ld, KC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

int t5_pick_bn(int N) {
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    return 256;
}

// NT / NN: the persistent grid runs tiles_m x tiles_n work items in waves of one per SM.  With few row tiles (the 16 x 16
// bottleneck stages: M = 4096) the widest column tile leaves most SMs idle -- N = 512 at BN = 256 is 64 items on 148 SMs --
// so take the narrower tile when that shortens the critical path (per-item cost ~ the operand bytes per chunk, A + B).
int t5_pick_bn_mn(int M, int N) {
    int best = t5_pick_bn(N);
    if (best <= 64) return best;
    const int sms = uwr_sm_count();
    const long long tm = uwr_cdiv(M, TM);
    long long best_cost = ((tm * uwr_cdiv(N, best) + sms - 1) / sms) * (TM + best);
    for (int bn = best / 2; bn >= 64; bn /= 2) {
        const long long cost = ((tm * uwr_cdiv(N, bn) + sms - 1) / sms) * (TM + bn);
        if (cost * 10 < best_cost * 9) {   // at least 10 % shorter
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

struct T5Split {
    int splits, chunks_per_split;
};
T5Split t5_plan(int M, int N, int Kc, int lay, int bn) {
    const int chunks = uwr_cdiv(Kc, KC);
    T5Split sp{1, chunks};
    if (lay != LAY_TN) return sp;
    const long long tiles = (long long)uwr_cdiv(M, TM) * uwr_cdiv(N, bn);
    const int sms = uwr_sm_count();
    const int max_splits = chunks / 8 > 0 ? chunks / 8 : 1;  // >= 256 contraction rows per split
    // the persistent grid runs tiles * splits work items in waves of one per SM: pick the split count with the fewest
    // chunk-steps on the critical path.  (Rounding the split count UP, as this did before, made 150 items out of 3 tiles x
    // 50 splits -- two waves on 148 SMs, i.e. twice the time of 3 x 49.)
    int best = 1;
    long long best_cost = -1;
    int limit = (int)(2 * sms / tiles) + 1;
    if (limit > max_splits) limit = max_splits;
    for (int s = 1; s <= limit; ++s) {
        const int cps = uwr_cdiv(chunks, s), real = uwr_cdiv(chunks, cps);
        const long long waves = (tiles * real + sms - 1) / sms;
        const long long cost = waves * (cps + 6);             // + a fixed per-item cost (pipeline fill, epilogue)
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = s;
        }
    }
    sp.chunks_per_split = uwr_cdiv(chunks, best);
    sp.splits = uwr_cdiv(chunks, sp.chunks_per_split);
    return sp;
}

template <int BN, int LAY, int EPI, int HF = 0, int CL = 1, int CV = 0>
int t5_launch(const CUtensorMap& ma, const CUtensorMap& mb, const T5Params& p, cudaStream_t stream) {
    constexpr int smem = t5_smem_bytes<BN>();
    auto kern = gemm_tcgen05_kernel<BN, LAY, EPI, HF, CL, CV>;
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int items = ((p.tiles_m + CL - 1) / CL) * p.tiles_n * p.splits;   // per cluster
    int clusters = uwr_sm_count() / CL;
    if (clusters > items) clusters = items;
    if (CL == 1) {
        UWR_CUDA(uwr_launch_pdl(kern, dim3(clusters), dim3(T5_THREADS), smem, stream, ma, mb, p));
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * CL);
        cfg.blockDim = dim3(T5_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = uwr_pdl_enabled();
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        UWR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
    }
    UWR_CHECK_LAUNCH("gemm_tcgen05_kernel");
    return 0;
}

template <int BN, int LAY>
int t5_dispatch_epi(const CUtensorMap& ma, const CUtensorMap& mb, const T5Params& p, int hf, int cl, cudaStream_t stream) {
    // fp16 storage is instantiated for the two LeFF uses only: linear1 forward (NT, C half) and the linear2 data
    // gradient with the gelu' multiplier (NN, R half)
    if (hf == 1) return t5_launch<BN, LAY_NT, UWR_EPI_NONE, 1>(ma, mb, p, stream);
    if (hf == 2) return t5_launch<BN, LAY_NN, UWR_EPI_MUL, 2>(ma, mb, p, stream);
    if (cl == 2) {   // cluster pair with B-tile multicast: the tensor-bound shapes (BN >= 128 only)
        if constexpr (BN >= 128) {
            switch (p.epilogue) {
                case UWR_EPI_RESID: return t5_launch<BN, LAY, UWR_EPI_RESID, 0, 2>(ma, mb, p, stream);
                case UWR_EPI_MUL: return t5_launch<BN, LAY, UWR_EPI_MUL, 0, 2>(ma, mb, p, stream);
                case UWR_EPI_NONE: return t5_launch<BN, LAY, UWR_EPI_NONE, 0, 2>(ma, mb, p, stream);
                default: break;
            }
        }
    }
    switch (p.epilogue) {
        case UWR_EPI_RESID: return t5_launch<BN, LAY, UWR_EPI_RESID>(ma, mb, p, stream);
        case UWR_EPI_MUL: return t5_launch<BN, LAY, UWR_EPI_MUL>(ma, mb, p, stream);
        case UWR_EPI_MUL_DGELU: return t5_launch<BN, LAY, UWR_EPI_MUL_DGELU>(ma, mb, p, stream);
        default: return t5_launch<BN, LAY, UWR_EPI_NONE>(ma, mb, p, stream);
    }
}

template <int BN>
int t5_dispatch(int lay, const CUtensorMap& ma, const CUtensorMap& mb, const T5Params& p, int hf, int cl, cudaStream_t stream) {
    switch (lay) {
        case LAY_NT: return t5_dispatch_epi<BN, LAY_NT>(ma, mb, p, hf, cl, stream);
        case LAY_NN: return t5_dispatch_epi<BN, LAY_NN>(ma, mb, p, hf, cl, stream);
        default:
            if constexpr (BN >= 128) {
                if (cl == 2) return t5_launch<BN, LAY_TN, UWR_EPI_NONE, 0, 2>(ma, mb, p, stream);
            }
            return t5_launch<BN, LAY_TN, UWR_EPI_NONE>(ma, mb, p, stream);
    }
}

// Cluster pairs halve the L2 reads of the B tile: at least two row tiles to pair, a B tile wide enough to split, high
// arithmetic intensity (flop per algorithmic byte).  Measured in isolation (profiles/r2_kernel_bench.txt, t5big) the
// weight-gradient (TN) shapes gain 6..8 % and NT / NN nothing -- those are bound by each SM's own operand ingress (48 KB
// per 512 MMA cycles), which multicast does not reduce -- but inside the training step, where neighbouring kernels
// compete for L2, pairing every eligible layout is the fastest setting (same box, B = 16: off 445.2, TN only 446.6,
// all layouts 449.7 img/s).  0 = off, 1 = auto (every eligible shape), 2 = TN layout only.
int g_t5_cluster = 1;
int t5_pick_cluster(const uwr_gemm_desc* d, int bn, int hf) {
    if (!g_t5_cluster || hf || bn < 128 || d->epilogue == UWR_EPI_MUL_DGELU) return 1;
    if (g_t5_cluster == 2 && !d->a_km) return 1;
    if (uwr_cdiv(d->M, TM) < 2) return 1;
    const double flops = 2.0 * d->M * d->N * d->K;
    const double bytes = 4.0 * ((double)d->M * d->K + (double)d->K * d->N + (double)d->M * d->N);
    return flops / bytes >= 96.0 ? 2 : 1;
}

}  // namespace

// 0 = never pair CTAs, 1 = auto (default: every eligible shape), 2 = weight-gradient (TN) layout only
extern "C" int uwr_set_gemm_cluster(int mode) {
    if (mode < 0 || mode > 2) return -1;
    g_t5_cluster = mode;
    return 0;
}

extern "C" size_t uwr_gemm_tcgen05_workspace_bytes(int M, int N, int K, int a_km) {
    if (!a_km) return 0;
    const T5Split sp = t5_plan(M, N, K, LAY_TN, t5_pick_bn(N));
    return sp.splits > 1 ? (size_t)sp.splits * M * N * sizeof(float) : 0;
}

// 1 if the shape/layout/alignment of `d` is served by the tcgen05 path (the caller falls back to
// uwr_gemm_tf32 otherwise, e.g. segmented weights, odd widths, colsum / k-scale requests).
extern "C" int uwr_gemm_tcgen05_supported(const uwr_gemm_desc* d) {
    if (!d || d->B2 || d->colsum) return 0;
    if (d->a_km && (d->b_nk || d->rowscale || d->epilogue != UWR_EPI_NONE || d->bias)) return 0;
    if (d->K % 4 || d->N % 4 || d->lda % 4 || d->ldb % 4 || d->ldc % 4) return 0;
    if (((uintptr_t)d->A | (uintptr_t)d->B | (uintptr_t)d->C) % 16) return 0;
    if (d->epilogue != UWR_EPI_NONE && (!d->R || d->ldr % 4 || (uintptr_t)d->R % 16)) return 0;
    // fp16 storage: C half for a plain NT product (linear1 forward), R half for the NN product with the stored
    // gelu' multiplier (linear2 data gradient)
    if (d->c_half && !(d->b_nk && !d->a_km && d->epilogue == UWR_EPI_NONE && !d->r_half && d->N % 8 == 0 && d->ldc % 8 == 0))
        return 0;
    if (d->r_half && !(!d->b_nk && !d->a_km && d->epilogue == UWR_EPI_MUL && !d->c_half)) return 0;
    if (d->bias && (uintptr_t)d->bias % 16) return 0;
    // MN-major operands are fetched in 32-float groups; a ragged last group is zero-filled by TMA (out-of-bounds
    // elements of a box read as 0) and masked in the epilogue, so any width in multiples of 4 (16 B rows) is served
    if (d->a_km && d->M % 4) return 0;
    if (d->M < 1 || d->N < 8 || d->K < 8) return 0;
    return 1;
}

extern "C" int uwr_gemm_tcgen05(const uwr_gemm_desc* d, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(uwr_gemm_tcgen05_supported(d), "uwr_gemm_tcgen05: unsupported shape/layout (use uwr_gemm_tf32)");
    const int lay = d->a_km ? LAY_TN : (d->b_nk ? LAY_NT : LAY_NN);
    const int hf = (d->c_half ? 1 : 0) | (d->r_half ? 2 : 0);
    const int bn = (lay == LAY_TN || hf) ? t5_pick_bn(d->N) : t5_pick_bn_mn(d->M, d->N);
    const T5Split sp = t5_plan(d->M, d->N, d->K, lay, bn);
    const int cl = t5_pick_cluster(d, bn, hf);

    CUtensorMap ma, mb;
    int rc;
    if (lay == LAY_TN) {
        // A stored [K][M] (width M), B stored [K][N] (width N)
        if ((rc = encode_mn(&ma, d->A, d->M, d->K, d->lda))) return rc;
        if ((rc = encode_mn(&mb, d->B, d->N, d->K, d->ldb))) return rc;
    } else {
        if ((rc = encode_2d(&ma, d->A, d->K, d->M, d->lda, TM))) return rc;
        if (lay == LAY_NT) {   // K-major B: with a cluster pair each CTA's box is its half of the tile's rows
            if ((rc = encode_2d(&mb, d->B, d->K, d->N, d->ldb, bn / cl))) return rc;
        } else {
            if ((rc = encode_mn(&mb, d->B, d->N, d->K, d->ldb))) return rc;
        }
    }

    T5Params p{};
    p.C = d->C; p.ldc = d->ldc; p.M = d->M; p.N = d->N;
    p.chunks = uwr_cdiv(d->K, KC);
    p.chunks_per_split = sp.chunks_per_split; p.splits = sp.splits; p.split_stride = 0;
    p.tiles_m = uwr_cdiv(d->M, TM); p.tiles_n = uwr_cdiv(d->N, bn);
    p.bias = d->bias; p.R = d->R; p.ldr = d->ldr;
    p.rowscale = d->rowscale; p.rows_per_group = d->rows_per_group > 0 ? d->rows_per_group : 1;
    p.epilogue = d->epilogue;
    p.round_out = d->round_out;
    if (sp.splits > 1) {
        const size_t need = (size_t)sp.splits * d->M * d->N * sizeof(float);
        UWR_REQUIRE(d->workspace && d->workspace_bytes >= need, "uwr_gemm_tcgen05: workspace too small (%zu < %zu)",
                    d->workspace_bytes, need);
        UWR_REQUIRE(d->ldc == d->N, "uwr_gemm_tcgen05: split contraction needs a dense C");
        p.C = d->workspace; p.ldc = d->N; p.split_stride = (long long)d->M * d->N;
    }
    switch (bn) {
        case 32: rc = t5_dispatch<32>(lay, ma, mb, p, hf, cl, stream); break;
        case 64: rc = t5_dispatch<64>(lay, ma, mb, p, hf, cl, stream); break;
        case 128: rc = t5_dispatch<128>(lay, ma, mb, p, hf, cl, stream); break;
        default: rc = t5_dispatch<256>(lay, ma, mb, p, hf, cl, stream); break;
    }
    if (rc) return rc;
    if (sp.splits > 1) {
        const long long n = (long long)d->M * d->N;
        int blocks = (int)((n / 4 + 31) / 32);
        if (blocks > 8 * uwr_sm_count()) blocks = 8 * uwr_sm_count();
        UWR_CUDA(uwr_launch_pdl(t5_splitk_reduce_kernel, dim3(blocks), dim3(32 * RED_GROUPS), 0, stream,
                                (const float*)d->workspace, d->C, n, n, sp.splits));
        UWR_CHECK_LAUNCH("t5_splitk_reduce_kernel");
    }
    return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// Convolutions on token (NHWC) tensors as implicit GEMMs: the im2col matrix is a *view* that TMA materialises tile by
// tile straight into shared memory (T5Conv above), never in HBM.  Replaces im2col_* + uwr_gemm_tcgen05 for
//   3x3 stride 1 pad 1   (block.py:42-153 Down/Upsample + in/out convs, SpectralTransformer.py:133-159)
//   4x4 stride 2 pad 1   (AST.py:408-424 Downsample)
// mode 0: y = im2col(x) w^T + bias (forward; also the stride-1 data gradient, called with flipped weights);
// mode 1: dw = dy^T im2col(x) (weight gradient, contraction over the output pixels, split across CTAs).
namespace {

struct ConvGeom {
    int OH, OW, taps, bw, bh, pw, ph;
    bool ok_rows, ok_patch;   // the 128-pixel row box (mode 0) / the 32-pixel contraction patch (modes 1, 2) tile the image
};
bool conv_geom(const uwr_convgemm_desc* d, ConvGeom& g) {
    if (d->kh < 1 || d->kw < 1 || d->kh > 7 || d->kw > 7 || d->stride < 1 || d->stride > 2 || d->pad < 0 || d->pad > 3) return false;
    g.OH = (d->H + 2 * d->pad - d->kh) / d->stride + 1;
    g.OW = (d->W + 2 * d->pad - d->kw) / d->stride + 1;
    g.taps = d->kh * d->kw;
    if (g.OH < 1 || g.OW < 1) return false;
    // mode 0: one box = 128 output pixels = bw x bh (part of a row, or whole rows); mode 1: 32 output pixels = pw x ph
    g.bw = g.OW < TM ? g.OW : TM;  g.bh = TM / g.bw;
    g.pw = g.OW < KC ? g.OW : KC;  g.ph = KC / g.pw;
    // OW a power of two below the box, or a multiple of it; boxes never straddle rows / images; TMA box limit 256
    g.ok_rows = g.bw * g.bh == TM && g.OW % g.bw == 0 && g.OH % g.bh == 0 && g.bw * d->stride <= 256 && g.bh * d->stride <= 256;
    g.ok_patch = g.pw * g.ph == KC && g.OW % g.pw == 0 && g.OH % g.ph == 0;
    return true;
}
// (C, W, H, B) view of the token matrix; the box covers bw x bh output pixels of one tap (element strides = conv stride)
int encode_conv(CUtensorMap* m, const uwr_convgemm_desc* d, int bw, int bh, CUtensorMapSwizzle swz) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->ld_x * 4, (cuuint64_t)d->ld_x * 4 * d->W, (cuuint64_t)d->ld_x * 4 * d->W * d->H};
    cuuint32_t box[4] = {KC, (cuuint32_t)((bw - 1) * d->stride + 1), (cuuint32_t)((bh - 1) * d->stride + 1), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        uwr_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return -3;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)d->x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        uwr_set_error("cuTensorMapEncodeTiled(conv 4d) failed: %d (C %d W %d H %d B %d ld %lld box %u x %u stride %d)", (int)r,
                      d->Cin, d->W, d->H, d->B, d->ld_x, box[1], box[2], d->stride);
        return -3;
    }
    return 0;
}

template <int BN>
int conv_dispatch(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const T5Params& p, cudaStream_t stream) {
    if (mode == 0) return t5_launch<BN, LAY_NT, UWR_EPI_NONE, 0, 1, 1>(ma, mb, p, stream);
    if (mode == 2) return t5_launch<BN, LAY_TN, UWR_EPI_NONE, 0, 1, 3>(ma, mb, p, stream);
    return t5_launch<BN, LAY_TN, UWR_EPI_NONE, 0, 1, 2>(ma, mb, p, stream);
}

}  // namespace

extern "C" int uwr_convgemm_tcgen05_supported(const uwr_convgemm_desc* d) {
    ConvGeom g;
    if (!d || !d->x || d->Cin % 32 || d->ld_x % 4 || (uintptr_t)d->x % 16 || !conv_geom(d, g)) return 0;
    if (d->B < 1 || d->Cout < 8 || d->Cout % 4) return 0;
    const long long pixels = (long long)d->B * g.OH * g.OW;
    if (d->mode == 0) {
        if (!d->w || !d->y || d->ld_y % 4 || ((uintptr_t)d->w | (uintptr_t)d->y) % 16) return 0;
        if (d->bias && (uintptr_t)d->bias % 16) return 0;
        return g.ok_rows && (g.OH * g.OW) % TM == 0 && pixels < (1ll << 31);
    }
    if (d->mode == 1 || d->mode == 2) {
        if (!d->dy || !d->dw || d->ld_dy % 4 || ((uintptr_t)d->dy | (uintptr_t)d->dw) % 16) return 0;
        return g.ok_patch && (g.OH * g.OW) % KC == 0 && pixels < (1ll << 31);
    }
    return 0;
}

extern "C" size_t uwr_convgemm_tcgen05_workspace_bytes(const uwr_convgemm_desc* d) {
    ConvGeom g;
    if (!d || (d->mode != 1 && d->mode != 2) || !conv_geom(d, g)) return 0;
    const int KN = g.taps * d->Cin;
    const int M = d->mode == 1 ? d->Cout : KN, N = d->mode == 1 ? KN : d->Cout;
    const long long K = (long long)d->B * g.OH * g.OW;
    const T5Split sp = t5_plan(M, N, (int)K, LAY_TN, t5_pick_bn(N));
    return sp.splits > 1 ? (size_t)sp.splits * M * N * sizeof(float) : 0;
}

extern "C" int uwr_convgemm_tcgen05(const uwr_convgemm_desc* d, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(uwr_convgemm_tcgen05_supported(d), "uwr_convgemm_tcgen05: unsupported geometry / alignment");
    ConvGeom g;
    conv_geom(d, g);
    const long long pixels = (long long)d->B * g.OH * g.OW;
    const int KN = g.taps * d->Cin;   // im2col width
    T5Params p{};
    p.cv.cpt = d->Cin / KC; p.cv.kw = d->kw; p.cv.taps = g.taps; p.cv.pad = d->pad; p.cv.stride = d->stride;
    p.cv.ow = g.OW; p.cv.ohw = g.OH * g.OW; p.cv.cin = d->Cin; p.cv.nb = d->B; p.cv.pw = g.pw; p.cv.ph = g.ph;
    p.rows_per_group = 1;
    p.epilogue = UWR_EPI_NONE;
    CUtensorMap ma, mb;
    int rc, bn;
    T5Split sp{1, 0};
    if (d->mode == 0) {
        bn = t5_pick_bn(d->Cout);
        if ((rc = encode_conv(&ma, d, g.bw, g.bh, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = encode_2d(&mb, d->w, KN, d->Cout, KN, bn))) return rc;
        p.C = d->y; p.ldc = d->ld_y; p.M = (int)pixels; p.N = d->Cout;
        p.chunks = KN / KC; p.chunks_per_split = p.chunks; p.splits = 1;
        p.bias = d->bias; p.round_out = d->round_out;
    } else if (d->mode == 2) {
        // transposed weight gradient dw^T (KN, Cout) = im2col(x)^T dy: the im2col view is the (MN-major) A operand, so a
        // thin Cout costs a thin N tile instead of padding the 128-row M tile
        bn = t5_pick_bn(d->Cout);
        sp = t5_plan(KN, d->Cout, (int)pixels, LAY_TN, bn);
        if ((rc = encode_conv(&ma, d, g.pw, g.ph, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
        if ((rc = encode_mn(&mb, d->dy, d->Cout, pixels, d->ld_dy))) return rc;
        p.C = d->dw; p.ldc = d->Cout; p.M = KN; p.N = d->Cout;
        p.chunks = (int)(pixels / KC); p.chunks_per_split = sp.chunks_per_split; p.splits = sp.splits;
        if (sp.splits > 1) {
            const size_t need = (size_t)sp.splits * d->Cout * KN * sizeof(float);
            UWR_REQUIRE(d->workspace && d->workspace_bytes >= need, "uwr_convgemm_tcgen05: workspace too small (%zu < %zu)",
                        d->workspace_bytes, need);
            p.C = d->workspace; p.split_stride = (long long)d->Cout * KN;
        }
    } else {
        bn = t5_pick_bn(KN);
        sp = t5_plan(d->Cout, KN, (int)pixels, LAY_TN, bn);
        if ((rc = encode_mn(&ma, d->dy, d->Cout, pixels, d->ld_dy))) return rc;
        if ((rc = encode_conv(&mb, d, g.pw, g.ph, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
        p.C = d->dw; p.ldc = KN; p.M = d->Cout; p.N = KN;
        p.chunks = (int)(pixels / KC); p.chunks_per_split = sp.chunks_per_split; p.splits = sp.splits;
        if (sp.splits > 1) {
            const size_t need = (size_t)sp.splits * d->Cout * KN * sizeof(float);
            UWR_REQUIRE(d->workspace && d->workspace_bytes >= need, "uwr_convgemm_tcgen05: workspace too small (%zu < %zu)",
                        d->workspace_bytes, need);
            p.C = d->workspace; p.split_stride = (long long)d->Cout * KN;
        }
    }
    p.tiles_m = uwr_cdiv(p.M, TM); p.tiles_n = uwr_cdiv(p.N, bn);
    switch (bn) {
        case 32: rc = conv_dispatch<32>(d->mode, ma, mb, p, stream); break;
        case 64: rc = conv_dispatch<64>(d->mode, ma, mb, p, stream); break;
        case 128: rc = conv_dispatch<128>(d->mode, ma, mb, p, stream); break;
        default: rc = conv_dispatch<256>(d->mode, ma, mb, p, stream); break;
    }
    if (rc) return rc;
    if (d->mode != 0 && sp.splits > 1) {
        const long long n = (long long)d->Cout * KN;
        int blocks = (int)((n / 4 + 31) / 32);
        if (blocks > 8 * uwr_sm_count()) blocks = 8 * uwr_sm_count();
        UWR_CUDA(uwr_launch_pdl(t5_splitk_reduce_kernel, dim3(blocks), dim3(32 * RED_GROUPS), 0, stream,
                                (const float*)d->workspace, d->dw, n, n, sp.splits));
        UWR_CHECK_LAUNCH("t5_splitk_reduce_kernel");
    }
    return 0;
}
