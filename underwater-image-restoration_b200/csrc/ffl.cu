// Focal frequency loss (third-party focal_frequency_loss 0.3.0 as constructed at
// src/Losses/losses.py:48: loss_weight=1, alpha=1, patch_factor=1): value and gradient with
// shared-memory FFTs.
//
//   d = pred - truth (one S x S plane per (n, c));  D = FFT2_ortho(d)   (linearity: one FFT on d)
//   w = |D| / max_plane|D|   (NaN -> 0, clamp [0,1], treated as a constant)
//   loss = mean(w |D|^2) ;  dloss/dd = (2/n) Re(IFFT2_ortho(w * D))
//
// Row pass -> column pass (32-column tiles staged in shared memory, so every global access is
// coalesced) -> per-plane max -> weighting + inverse column pass -> inverse row pass.  Each 1-D
// transform is a radix-2 DIT FFT executed by one warp on shared memory (bit-reversed on load).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int FFL_WARPS = 8;
constexpr int COLS_PER_CTA = 32;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-place radix-2 DIT on x[0..S) (already bit-reversed), one warp; tw[k] = exp(-2 pi i k / S)
__device__ __forceinline__ void warp_fft(float2* x, const float2* tw, int S, int log2S, bool inverse, int lane) {
    for (int s = 1; s <= log2S; ++s) {
        const int m = 1 << s, half = m >> 1, tstep = S >> s;
        for (int b = lane; b < S / 2; b += 32) {
            const int j = b & (half - 1);
            const int i0 = ((b >> (s - 1)) << s) + j, i1 = i0 + half;
            float2 w = tw[j * tstep];
            if (inverse) w.y = -w.y;
            const float2 t = cmul(w, x[i1]);
            const float2 a = x[i0];
            x[i0] = make_float2(a.x + t.x, a.y + t.y);
            x[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void make_twiddles(float2* tw, int S) {
    for (int k = threadIdx.x; k < S / 2; k += blockDim.x) {
        float sn, cs;
        sincospif(-2.0f * (float)k / (float)S, &sn, &cs);
        tw[k] = make_float2(cs, sn);
    }
}

// rows, forward: one warp per row.  smem: tw[S/2] + FFL_WARPS * S complex
__global__ void __launch_bounds__(FFL_WARPS * 32) ffl_rows_fwd_kernel(const float* __restrict__ pred,
                                                                     const float* __restrict__ truth,
                                                                     float2* __restrict__ W1,
                                                                     unsigned* __restrict__ plane_max, int S,
                                                                     int log2S, float scale) {
    uwr_pdl_enter();
    extern __shared__ __align__(16) float2 sm2[];
    float2* tw = sm2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* x = sm2 + S / 2 + warp * S;
    make_twiddles(tw, S);
    const int plane = blockIdx.x, row = blockIdx.y * FFL_WARPS + warp;
    if (blockIdx.y == 0 && threadIdx.x == 0) plane_max[plane] = 0u;
    const long long base = ((long long)plane * S + row) * S;
    if (row < S)
        for (int c = lane; c < S; c += 32)
            x[__brev((unsigned)c) >> (32 - log2S)] = make_float2(pred[base + c] - truth[base + c], 0.f);
    __syncthreads();
    if (row >= S) return;
    warp_fft(x, tw, S, log2S, false, lane);
    for (int c = lane; c < S; c += 32) W1[base + c] = make_float2(x[c].x * scale, x[c].y * scale);
}

// columns: CTA = 32 columns x S rows of one plane.  smem: tw[S/2] + 32 * (S+1) complex.
// MODE 0: forward FFT, write D, update the plane maximum of |D|.
// MODE 1: read D, weight by |D|/max, accumulate the loss partial, inverse FFT, write back.
template <int MODE>
__global__ void __launch_bounds__(FFL_WARPS * 32) ffl_cols_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                                 unsigned* __restrict__ plane_max,
                                                                 float* __restrict__ partials, int S, int log2S,
                                                                 float scale) {
    uwr_pdl_enter();
    extern __shared__ __align__(16) float2 sm2[];
    __shared__ float red[FFL_WARPS];
    float2* tw = sm2;
    float2* tile = sm2 + S / 2;  // [32][S+1]
    const int pitch = S + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    make_twiddles(tw, S);
    const int plane = blockIdx.x, c0 = blockIdx.y * COLS_PER_CTA;
    const long long pbase = (long long)plane * S * S;
    float local = 0.f;
    float inv_max = 0.f;
    if (MODE == 1) {
        const float m = __uint_as_float(plane_max[plane]);
        inv_max = m > 0.f ? 1.0f / m : 0.f;  // max == 0 -> the reference's NaN weights become 0
    }
    for (int idx = threadIdx.x; idx < S * COLS_PER_CTA; idx += blockDim.x) {
        const int r = idx / COLS_PER_CTA, c = idx % COLS_PER_CTA;
        float2 v = in[pbase + (long long)r * S + c0 + c];
        if (MODE == 1) {
            const float mag2 = v.x * v.x + v.y * v.y;
            const float w = fminf(sqrtf(mag2) * inv_max, 1.0f);
            local += w * mag2;
            v = make_float2(v.x * w, v.y * w);
        }
        tile[c * pitch + (__brev((unsigned)r) >> (32 - log2S))] = v;
    }
    __syncthreads();
    for (int c = warp; c < COLS_PER_CTA; c += FFL_WARPS) warp_fft(tile + c * pitch, tw, S, log2S, MODE == 1, lane);
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * COLS_PER_CTA; idx += blockDim.x) {
        const int r = idx / COLS_PER_CTA, c = idx % COLS_PER_CTA;
        const float2 v = tile[c * pitch + r];
        const float2 o = make_float2(v.x * scale, v.y * scale);
        out[pbase + (long long)r * S + c0 + c] = o;
        if (MODE == 0) local = fmaxf(local, sqrtf(o.x * o.x + o.y * o.y));
    }
    local = MODE == 0 ? warp_max(local) : warp_sum(local);
    if (lane == 0) red[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = red[0];
        for (int w = 1; w < FFL_WARPS; ++w) a = MODE == 0 ? fmaxf(a, red[w]) : a + red[w];
        if (MODE == 0) atomicMax(plane_max + plane, __float_as_uint(a));  // |D| >= 0: uint order == float order
        else partials[plane * gridDim.y + blockIdx.y] = a;
    }
}

// rows, inverse: one warp per row, real part scaled into the gradient
__global__ void __launch_bounds__(FFL_WARPS * 32) ffl_rows_inv_kernel(const float2* __restrict__ W1,
                                                                     float* __restrict__ grad, int S, int log2S,
                                                                     float scale) {
    uwr_pdl_enter();
    extern __shared__ __align__(16) float2 sm2[];
    float2* tw = sm2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* x = sm2 + S / 2 + warp * S;
    make_twiddles(tw, S);
    const int plane = blockIdx.x, row = blockIdx.y * FFL_WARPS + warp;
    const long long base = ((long long)plane * S + row) * S;
    if (row < S)
        for (int c = lane; c < S; c += 32) x[__brev((unsigned)c) >> (32 - log2S)] = W1[base + c];
    __syncthreads();
    if (row >= S) return;
    warp_fft(x, tw, S, log2S, true, lane);
    for (int c = lane; c < S; c += 32) grad[base + c] = x[c].x * scale;
}

__global__ void ffl_final_kernel(const float* __restrict__ partials, int n, float inv_count, float* __restrict__ out) {
    uwr_pdl_enter();
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) a += partials[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 32; ++i) s += sh[i];
        out[0] = (float)(s * inv_count);
    }
}

int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace

extern "C" size_t uwr_ffl_workspace_bytes(int planes, int S) {
    const size_t plane_elems = (size_t)planes * S * S;
    return 2 * plane_elems * sizeof(float2) + (size_t)planes * sizeof(unsigned) +
           (size_t)planes * (S / COLS_PER_CTA) * sizeof(float) + 64;
}

extern "C" int uwr_ffl_loss(const float* pred, const float* truth, float* out, float* grad, float* workspace,
                            int planes, int S, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(pred && truth && out && workspace, "uwr_ffl_loss: null pointer");
    UWR_REQUIRE(S >= 64 && S <= 512 && (S & (S - 1)) == 0, "uwr_ffl_loss: S must be a power of two in [64, 512]");
    UWR_REQUIRE(planes > 0 && planes <= 65535, "uwr_ffl_loss: bad plane count");
    const int log2S = ilog2(S);
    const size_t plane_elems = (size_t)planes * S * S;
    float2* W1 = reinterpret_cast<float2*>(workspace);
    float2* W2 = W1 + plane_elems;
    unsigned* pmax = reinterpret_cast<unsigned*>(W2 + plane_elems);
    float* partials = reinterpret_cast<float*>(pmax + planes);
    const float sc = 1.0f / sqrtf((float)S);  // ortho: 1/sqrt(S) per 1-D pass
    const int row_smem = (S / 2 + FFL_WARPS * S) * (int)sizeof(float2);
    const int col_smem = (S / 2 + COLS_PER_CTA * (S + 1)) * (int)sizeof(float2);
    static bool configured = false;
    if (!configured) {
        UWR_CUDA(cudaFuncSetAttribute(ffl_cols_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        UWR_CUDA(cudaFuncSetAttribute(ffl_cols_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
        configured = true;
    }
    const dim3 rgrid(planes, uwr_cdiv(S, FFL_WARPS)), cgrid(planes, S / COLS_PER_CTA);
    (void)uwr_launch_pdl(ffl_rows_fwd_kernel, dim3(rgrid), dim3(FFL_WARPS * 32), row_smem, stream, pred, truth, W1, pmax, S, log2S, sc);
    UWR_CHECK_LAUNCH("ffl_rows_fwd_kernel");
    (void)uwr_launch_pdl(ffl_cols_kernel<0>, dim3(cgrid), dim3(FFL_WARPS * 32), col_smem, stream, W1, W2, pmax, partials, S, log2S, sc);
    UWR_CHECK_LAUNCH("ffl_cols_kernel<0>");
    (void)uwr_launch_pdl(ffl_cols_kernel<1>, dim3(cgrid), dim3(FFL_WARPS * 32), col_smem, stream, W2, W1, pmax, partials, S, log2S, sc);
    UWR_CHECK_LAUNCH("ffl_cols_kernel<1>");
    const double count = (double)planes * S * S;
    (void)uwr_launch_pdl(ffl_final_kernel, dim3(1), dim3(32), 0, stream, partials, planes * (S / COLS_PER_CTA), (float)(1.0 / count), out);
    UWR_CHECK_LAUNCH("ffl_final_kernel");
    if (grad) {
        (void)uwr_launch_pdl(ffl_rows_inv_kernel, dim3(rgrid), dim3(FFL_WARPS * 32), row_smem, stream, W1, grad, S, log2S, sc * (float)(2.0 / count));
        UWR_CHECK_LAUNCH("ffl_rows_inv_kernel");
    }
    return 0;
}
