// Shared-memory FFT passes for the frequency-domain mixing of the restoration transformers:
//
//   uwr_dft_hw_real : y = scale * Re(FFT2 over (H, W)) of a real token tensor (B, H, W, C)
//                     FDFP (src/model/block.py:532-556: fftn / ifftn over the spatial axes, .real)
//   uwr_dft_lc_real : y = scale * Re(FFT2 over (L = H*W, C)) of a real token tensor (B, L, C)
//                     EncoderBlock (src/model/model.py:72-88: fftn / ifftn over (tokens, channels), .real)
//   uwr_fft2_hw     : complex 2-D FFT over (H, W) of a (B, H, W, C) tensor, real or complex input,
//                     forward or inverse (SpectralTransformer.UpSample, SpectralTransformer.py:174-188)
//
// For a REAL input the real part of the inverse transform equals the real part of the forward one
// divided by the element count, and x -> Re(F x) is a symmetric linear map (C (x) C - S (x) S), so the
// same entry point with a different `scale` is the forward, the inverse and both of their backward
// passes.
//
// One pass = 1-D FFTs along one axis of the (B, Y, X, C) view.  A CTA stages a tile of 32 adjacent
// channels x N points in shared memory (every global access is a 128/256-byte row of channels), each
// warp runs radix-2 DIT butterflies on its lines in place, and the tile is written back with the
// pass's epilogue (scale, four-step twiddle, real part, transposed store).  The token-axis transform
// of length L = H*W (up to 65 536) is done as the classic four-step FFT on the (H, W) grid:
// length-H FFTs down the columns, twiddle exp(-2 pi i x k1 / L), length-W FFTs along the rows, and
// the result stored transposed (k = k1 + H k2).
// HBM-bound: each pass reads and writes the tile once; complex intermediates live in the workspace.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int FFT_WARPS = 8;
constexpr int FFT_CT = 32;  // channels (lines) per CTA tile

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-place radix-2 DIT on x[0..N) (bit-reversed on load), one warp; tw[k] = exp(-2 pi i k / N)
__device__ __forceinline__ void warp_fft_r2(float2* x, const float2* tw, int N, int log2N, bool inverse, int lane) {
    for (int s = 1; s <= log2N; ++s) {
        const int half = 1 << (s - 1), tstep = N >> s;
        for (int b = lane; b < N / 2; b += 32) {
            const int j = b & (half - 1);
            const int i0 = ((b >> (s - 1)) << s) + j, i1 = i0 + half;
            float2 w = tw[j * tstep];
            if (inverse) w.y = -w.y;
            const float2 t = cmulf(w, x[i1]);
            const float2 a = x[i0];
            x[i0] = make_float2(a.x + t.x, a.y + t.y);
            x[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncwarp();
    }
}

struct PassParams {
    const float* in;
    float* out;
    int in_complex;   // 0: real input, 1: interleaved complex
    int out_real;     // 0: interleaved complex output, 1: real part only
    int inverse;      // conjugate twiddles (no implicit 1/N: fold it into `scale`)
    int N, log2N;     // FFT length (axis extent)
    int NO;           // extent of the other spatial axis
    int C;            // channels (contiguous)
    long long in_sB, in_sO, in_sN;     // element strides (complex or real elements) of batch / other / FFT axis
    long long out_sB, out_sO, out_sN;
    int tw_L;         // four-step twiddle: multiply output k by exp(-/+ 2 pi i * o * k / tw_L); 0 = none
    float scale;
};

// grid = (channel groups, NO, B)
__global__ void __launch_bounds__(FFT_WARPS * 32) fft_pass_kernel(const PassParams p) {
    uwr_pdl_enter();
    extern __shared__ __align__(16) float2 fsm[];
    float2* tw = fsm;           // N twiddles (the radix-2 path uses the first N/2)
    float2* tile = fsm + p.N;   // [FFT_CT][N + 1] (+ a second tile for the direct-DFT path)
    const int N = p.N, pitch = N + 1;
    const bool radix2 = p.log2N >= 0;  // otherwise: N is not a power of two -> direct O(N^2) DFT per line
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
        tw[k] = make_float2(cs, sn);
    }
    const int c0 = blockIdx.x * FFT_CT, o = blockIdx.y, b = blockIdx.z;
    const int nc = min(FFT_CT, p.C - c0);
    // four-step twiddles of this tile (fixed o): tw4[k] = exp(-/+ 2 pi i * o * k / tw_L), once per CTA
    float2* tw4 = tile + (radix2 ? 1 : 2) * FFT_CT * pitch;
    if (p.tw_L)
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            const unsigned m = ((unsigned)o * (unsigned)k) & (unsigned)(p.tw_L - 1);  // tw_L is a power of two
            float sn, cs;
            sincospif((p.inverse ? 2.0f : -2.0f) * (float)m / (float)p.tw_L, &sn, &cs);
            tw4[k] = make_float2(cs, sn);
        }
    const long long ibase = (long long)b * p.in_sB + (long long)o * p.in_sO + c0;
    for (int idx = threadIdx.x; idx < N * FFT_CT; idx += blockDim.x) {
        const int n = idx >> 5, c = idx & 31;
        float2 v = make_float2(0.f, 0.f);
        if (c < nc) {
            const long long a = ibase + (long long)n * p.in_sN + c;
            if (p.in_complex) v = reinterpret_cast<const float2*>(p.in)[a];
            else v.x = p.in[a];
        }
        tile[c * pitch + (radix2 ? (int)(__brev((unsigned)n) >> (32 - p.log2N)) : n)] = v;
    }
    __syncthreads();
    if (radix2) {
        for (int c = warp; c < nc; c += FFT_WARPS) warp_fft_r2(tile + c * pitch, tw, N, p.log2N, p.inverse != 0, lane);
    } else {
        float2* res = tile + FFT_CT * pitch;
        for (int c = warp; c < nc; c += FFT_WARPS) {
            const float2* x = tile + c * pitch;
            for (int k = lane; k < N; k += 32) {
                float2 acc = make_float2(0.f, 0.f);
                int m = 0;  // (n * k) mod N
                for (int n = 0; n < N; ++n) {
                    float2 w = tw[m];
                    if (p.inverse) w.y = -w.y;
                    const float2 t = cmulf(w, x[n]);
                    acc.x += t.x;
                    acc.y += t.y;
                    m += k;
                    if (m >= N) m -= N;
                }
                res[c * pitch + k] = acc;
            }
        }
        tile = res;
    }
    __syncthreads();
    const long long obase = (long long)b * p.out_sB + (long long)o * p.out_sO + c0;
    for (int idx = threadIdx.x; idx < N * FFT_CT; idx += blockDim.x) {
        const int k = idx >> 5, c = idx & 31;
        if (c >= nc) continue;
        float2 v = tile[c * pitch + k];
        if (p.tw_L) v = cmulf(v, tw4[k]);
        const long long a = obase + (long long)k * p.out_sN + c;
        if (p.out_real) p.out[a] = v.x * p.scale;
        else reinterpret_cast<float2*>(p.out)[a] = make_float2(v.x * p.scale, v.y * p.scale);
    }
}

// FFT along the contiguous (channel) axis: one warp per row of N elements
__global__ void __launch_bounds__(FFT_WARPS * 32) fft_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                 int in_complex, int out_real, int inverse,
                                                                 long long rows, int N, int log2N, float scale) {
    uwr_pdl_enter();
    extern __shared__ __align__(16) float2 fsm[];
    float2* tw = fsm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* x = fsm + N / 2 + warp * N;
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
        tw[k] = make_float2(cs, sn);
    }
    __syncthreads();
    for (long long r = (long long)blockIdx.x * FFT_WARPS + warp; r < rows; r += (long long)gridDim.x * FFT_WARPS) {
        for (int c = lane; c < N; c += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (in_complex) v = reinterpret_cast<const float2*>(in)[r * N + c];
            else v.x = in[r * N + c];
            x[__brev((unsigned)c) >> (32 - log2N)] = v;
        }
        __syncwarp();
        warp_fft_r2(x, tw, N, log2N, inverse != 0, lane);
        for (int c = lane; c < N; c += 32) {
            if (out_real) out[r * N + c] = x[c].x * scale;
            else reinterpret_cast<float2*>(out)[r * N + c] = make_float2(x[c].x * scale, x[c].y * scale);
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------
// Register two-step variant (N = R1 * R2 <= 512): lanes <-> channels, every thread owns whole
// R-point sub-transforms in registers, so a tile makes ONE round trip through shared memory (the
// R1 x R2 transposition) instead of log2(N) radix-2 stages, and both global accesses are direct,
// fully coalesced rows of 32 channels.
//   step A: for each n2: R1-point FFT over n1 of x[R2*n1 + n2], times W_N^(n2*k1)  -> smem[k1][n2]
//   step B: for each k1: R2-point FFT over n2                                      -> X[k1 + R1*k2]

// W_32^k = exp(-2 pi i k / 32), k < 16 (compile-time after unrolling)
__device__ __forceinline__ float2 tw32(int k) {
    switch (k) {
        case 0: return make_float2(1.000000000f, -0.000000000f);
        case 1: return make_float2(0.980785280f, -0.195090322f);
        case 2: return make_float2(0.923879533f, -0.382683432f);
        case 3: return make_float2(0.831469612f, -0.555570233f);
        case 4: return make_float2(0.707106781f, -0.707106781f);
        case 5: return make_float2(0.555570233f, -0.831469612f);
        case 6: return make_float2(0.382683432f, -0.923879533f);
        case 7: return make_float2(0.195090322f, -0.980785280f);
        case 8: return make_float2(0.000000000f, -1.000000000f);
        case 9: return make_float2(-0.195090322f, -0.980785280f);
        case 10: return make_float2(-0.382683432f, -0.923879533f);
        case 11: return make_float2(-0.555570233f, -0.831469612f);
        case 12: return make_float2(-0.707106781f, -0.707106781f);
        case 13: return make_float2(-0.831469612f, -0.555570233f);
        case 14: return make_float2(-0.923879533f, -0.382683432f);
        case 15: return make_float2(-0.980785280f, -0.195090322f);
    }
    return make_float2(1.f, 0.f);
}

// in-register radix-2 DIF; natural-order output k is v[bit_reverse(k)]
template <int R, bool INV>
__device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
#pragma unroll
    for (int span = R / 2; span >= 1; span >>= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * span) {
#pragma unroll
            for (int j = 0; j < span; ++j) {
                const float2 a = v[blk + j], b = v[blk + j + span];
                v[blk + j] = make_float2(a.x + b.x, a.y + b.y);
                const float2 d = make_float2(a.x - b.x, a.y - b.y);
                if (j == 0) {
                    v[blk + j + span] = d;
                } else {
                    float2 w = tw32(j * (16 / span));  // W_(2 span)^j
                    if (INV) w.y = -w.y;
                    v[blk + j + span] = cmulf(d, w);
                }
            }
        }
    }
}

template <int R>
__host__ __device__ constexpr int brev_r(int k) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (k & 1);
        k >>= 1;
    }
    return r;
}

template <int R1, int R2, bool INV>
__global__ void __launch_bounds__(FFT_WARPS * 32) fft_pass2_kernel(const PassParams p) {
    uwr_pdl_enter();
    constexpr int N = R1 * R2;
    extern __shared__ __align__(16) float2 fsm[];
    float2* buf = fsm;                // [N][32]
    float2* twN = fsm + N * FFT_CT;   // W_N^m
    float2* tw4 = twN + N;            // four-step twiddles of this tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * FFT_CT, o = blockIdx.y, b = blockIdx.z;
    const bool cok = c0 + lane < p.C;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        float sn, cs;
        sincospif((INV ? 2.0f : -2.0f) * (float)k / (float)N, &sn, &cs);
        twN[k] = make_float2(cs, sn);
        if (p.tw_L) {
            const unsigned m = ((unsigned)o * (unsigned)k) & (unsigned)(p.tw_L - 1);
            sincospif((INV ? 2.0f : -2.0f) * (float)m / (float)p.tw_L, &sn, &cs);
            tw4[k] = make_float2(cs, sn);
        }
    }
    __syncthreads();
    const long long ibase = (long long)b * p.in_sB + (long long)o * p.in_sO + c0 + lane;
    for (int n2 = warp; n2 < R2; n2 += FFT_WARPS) {
        float2 v[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) {
            v[n1] = make_float2(0.f, 0.f);
            if (cok) {
                const long long a = ibase + (long long)(R2 * n1 + n2) * p.in_sN;
                if (p.in_complex) v[n1] = reinterpret_cast<const float2*>(p.in)[a];
                else v[n1].x = p.in[a];
            }
        }
        fft_reg<R1, INV>(v);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 val = v[brev_r<R1>(k1)];
            if (k1 > 0) val = cmulf(val, twN[n2 * k1]);  // n2 * k1 < N
            buf[(k1 * R2 + n2) * FFT_CT + lane] = val;
        }
    }
    __syncthreads();
    const long long obase = (long long)b * p.out_sB + (long long)o * p.out_sO + c0 + lane;
    for (int k1 = warp; k1 < R1; k1 += FFT_WARPS) {
        float2 u[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) u[n2] = buf[(k1 * R2 + n2) * FFT_CT + lane];
        fft_reg<R2, INV>(u);
        if (cok) {
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int k = k1 + R1 * k2;
                float2 val = u[brev_r<R2>(k2)];
                if (p.tw_L) val = cmulf(val, tw4[k]);
                const long long a = obase + (long long)k * p.out_sN;
                if (p.out_real) p.out[a] = val.x * p.scale;
                else reinterpret_cast<float2*>(p.out)[a] = make_float2(val.x * p.scale, val.y * p.scale);
            }
        }
    }
}

template <int R1, int R2>
int launch_pass2(const PassParams& p, int B, cudaStream_t stream) {
    constexpr int N = R1 * R2;
    constexpr int smem = (N * FFT_CT + 2 * N) * (int)sizeof(float2);
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(fft_pass2_kernel<R1, R2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        UWR_CUDA(cudaFuncSetAttribute(fft_pass2_kernel<R1, R2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const dim3 grid(uwr_cdiv(p.C, FFT_CT), p.NO, B);
    if (p.inverse) (void)uwr_launch_pdl(fft_pass2_kernel<R1, R2, true>, dim3(grid), dim3(FFT_WARPS * 32), smem, stream, p);
    else (void)uwr_launch_pdl(fft_pass2_kernel<R1, R2, false>, dim3(grid), dim3(FFT_WARPS * 32), smem, stream, p);
    UWR_CHECK_LAUNCH("fft_pass2_kernel");
    return 0;
}

int ilog2i(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}
bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int launch_pass(PassParams& p, int B, cudaStream_t stream) {
    UWR_REQUIRE(p.N >= 1 && (pow2(p.N) ? p.N <= 1024 : p.N <= 384),
                "fft pass: length %d unsupported (powers of two <= 1024, other lengths <= 384)", p.N);
    UWR_REQUIRE(B > 0 && B <= 65535 && p.NO > 0 && p.NO <= 65535, "fft pass: bad batch / extent");
    switch (p.N) {  // register two-step kernels
        case 16: return launch_pass2<4, 4>(p, B, stream);
        case 32: return launch_pass2<8, 4>(p, B, stream);
        case 64: return launch_pass2<8, 8>(p, B, stream);
        case 128: return launch_pass2<16, 8>(p, B, stream);
        case 256: return launch_pass2<16, 16>(p, B, stream);
        case 512: return launch_pass2<32, 16>(p, B, stream);
        default: break;
    }
    p.log2N = pow2(p.N) ? ilog2i(p.N) : -1;
    const int smem = (2 * p.N + (pow2(p.N) ? 1 : 2) * FFT_CT * (p.N + 1)) * (int)sizeof(float2);
    static int configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        UWR_CUDA(cudaFuncSetAttribute(fft_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    (void)uwr_launch_pdl(fft_pass_kernel, dim3(dim3(uwr_cdiv(p.C, FFT_CT), p.NO, B)), dim3(FFT_WARPS * 32), smem, stream, p);
    UWR_CHECK_LAUNCH("fft_pass_kernel");
    return 0;
}

int launch_rows(const float* in, float* out, int in_complex, int out_real, int inverse, long long rows, int N,
                float scale, cudaStream_t stream) {
    UWR_REQUIRE(pow2(N) && N >= 2 && N <= 1024, "fft rows: length %d must be a power of two in [2, 1024]", N);
    const int smem = (N / 2 + FFT_WARPS * N) * (int)sizeof(float2);
    static int configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        UWR_CUDA(cudaFuncSetAttribute(fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    long long blocks = (rows + FFT_WARPS - 1) / FFT_WARPS;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    (void)uwr_launch_pdl(fft_rows_kernel, dim3((unsigned)blocks), dim3(FFT_WARPS * 32), smem, stream, in, out, in_complex, out_real, inverse, rows, N,
                                                                       ilog2i(N), scale);
    UWR_CHECK_LAUNCH("fft_rows_kernel");
    return 0;
}

}  // namespace

extern "C" size_t uwr_dft_workspace_bytes(int B, int H, int W, int C) {
    return 2 * (size_t)B * H * W * C * sizeof(float2);  // two complex intermediates
}

// complex 2-D FFT over (H, W) of (B, H, W, C); in_complex / inverse as named; out is interleaved complex
extern "C" int uwr_fft2_hw(const float* in, float* out, float* workspace, int B, int H, int W, int C, int in_complex,
                           int inverse, float scale, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(in && out && workspace, "uwr_fft2_hw: null pointer");
    PassParams p{};
    // along X (W): line = (b, y); other axis = y
    p.in = in; p.out = workspace; p.in_complex = in_complex; p.out_real = 0; p.inverse = inverse;
    p.N = W; p.NO = H; p.C = C;
    p.in_sB = (long long)H * W * C; p.in_sO = (long long)W * C; p.in_sN = C;
    p.out_sB = p.in_sB; p.out_sO = p.in_sO; p.out_sN = C;
    p.tw_L = 0; p.scale = 1.0f;
    if (int rc = launch_pass(p, B, stream)) return rc;
    // along Y (H): line = (b, x)
    p.in = workspace; p.out = out; p.in_complex = 1; p.out_real = 0;
    p.N = H; p.NO = W;
    p.in_sO = C; p.in_sN = (long long)W * C; p.out_sO = C; p.out_sN = (long long)W * C;
    p.scale = scale;
    return launch_pass(p, B, stream);
}

extern "C" int uwr_dft_hw_real(const float* x, float* y, float* workspace, int B, int H, int W, int C, float scale,
                               uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && y && workspace, "uwr_dft_hw_real: null pointer");
    PassParams p{};
    p.in = x; p.out = workspace; p.in_complex = 0; p.out_real = 0; p.inverse = 0;
    p.N = W; p.NO = H; p.C = C;
    p.in_sB = (long long)H * W * C; p.in_sO = (long long)W * C; p.in_sN = C;
    p.out_sB = p.in_sB; p.out_sO = p.in_sO; p.out_sN = C;
    p.tw_L = 0; p.scale = 1.0f;
    if (int rc = launch_pass(p, B, stream)) return rc;
    p.in = workspace; p.out = y; p.in_complex = 1; p.out_real = 1;
    p.N = H; p.NO = W;
    p.in_sO = C; p.in_sN = (long long)W * C; p.out_sO = C; p.out_sN = (long long)W * C;
    p.scale = scale;
    return launch_pass(p, B, stream);
}

extern "C" int uwr_dft_lc_real(const float* x, float* y, float* workspace, int B, int H, int W, int C, float scale,
                               uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && y && workspace, "uwr_dft_lc_real: null pointer");
    UWR_REQUIRE(pow2(H) && pow2(W), "uwr_dft_lc_real: H and W must be powers of two");
    const long long L = (long long)H * W;
    UWR_REQUIRE(L <= (1LL << 30), "uwr_dft_lc_real: token axis too long");
    float* w1 = workspace;
    float* w2 = workspace + 2 * (size_t)B * L * C;
    // channel axis: real -> complex
    if (int rc = launch_rows(x, w1, 0, 0, 0, (long long)B * L, C, 1.0f, stream)) return rc;
    // token axis, four-step: l = y*W + x.  Step 1: length-H FFT over y for every (x, c), twiddle exp(-2 pi i x k1 / L)
    PassParams p{};
    p.in = w1; p.out = w2; p.in_complex = 1; p.out_real = 0; p.inverse = 0;
    p.N = H; p.NO = W; p.C = C;
    p.in_sB = L * C; p.in_sO = C; p.in_sN = (long long)W * C;
    p.out_sB = L * C; p.out_sO = C; p.out_sN = (long long)W * C;
    p.tw_L = (int)L; p.scale = 1.0f;
    if (int rc = launch_pass(p, B, stream)) return rc;
    // Step 2: length-W FFT over x for every (k1, c); frequency k = k1 + H*k2 -> token index k2*H + k1 (transposed store)
    p.in = w2; p.out = y; p.in_complex = 1; p.out_real = 1;
    p.N = W; p.NO = H;
    p.in_sO = (long long)W * C; p.in_sN = C;
    p.out_sO = C; p.out_sN = (long long)H * C;
    p.tw_L = 0; p.scale = scale;
    return launch_pass(p, B, stream);
}
