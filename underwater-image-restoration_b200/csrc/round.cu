// TF32 operand preparation for the tcgen05 GEMM path (kind::tf32 truncates, so operands are
// rounded to nearest once, where they are produced):
//   uwr_round_tf32_tensors : multi-tensor copy (+ optional rounding) -> rounded weight copies,
//                            also used to pack to_q|to_kv weights / biases contiguously
//   uwr_scale_round        : d_s = tf32(rowscale * d)  (DropPath-scaled residual-stream gradient,
//                            the A operand of the proj / linear2 data- and weight-gradient GEMMs)
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

__device__ __forceinline__ int rt_find(const long long* __restrict__ offsets, int n, long long i) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) round_tensors_kernel(const float* const* __restrict__ src,
                                                            float* const* __restrict__ dst,
                                                            const long long* __restrict__ offsets, int n,
                                                            long long total, int do_round) {
    uwr_pdl_enter();
    constexpr int CHUNK = 4096;
    for (long long c0 = (long long)blockIdx.x * CHUNK; c0 < total; c0 += (long long)gridDim.x * CHUNK) {
        const long long c1 = min(total, c0 + CHUNK);
        long long i = c0 + threadIdx.x;
        if (i >= c1) continue;
        int t = rt_find(offsets, n, i);
        for (; i < c1; i += 256) {
            while (i >= offsets[t + 1]) ++t;
            const float v = src[t][i - offsets[t]];
            dst[t][i - offsets[t]] = do_round ? tf32_round(v) : v;
        }
    }
}

__global__ void __launch_bounds__(256) scale_round_kernel(const float* __restrict__ src, long long ld_src,
                                                          float* __restrict__ dst, long long rows, int cols4,
                                                          const float* __restrict__ rowscale, int rpg,
                                                          int do_round) {
    uwr_pdl_enter();
    const long long total = rows * cols4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols4;
        const int c = (int)(i % cols4) * 4;
        float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
        const float s = rowscale ? __ldg(rowscale + r / rpg) : 1.f;
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        if (do_round) {
            v.x = tf32_round(v.x); v.y = tf32_round(v.y); v.z = tf32_round(v.z); v.w = tf32_round(v.w);
        }
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}

// scale + round + column sums: thread -> fixed float4 column group, rows strided; grid-level partials
__global__ void __launch_bounds__(256) scale_round_colsum_kernel(const float* __restrict__ src, long long ld_src,
                                                                 float* __restrict__ dst, long long rows, int cols4,
                                                                 const float* __restrict__ rowscale, int rpg,
                                                                 int do_round, float* __restrict__ partials) {
    uwr_pdl_enter();
    __shared__ float4 sh[256];
    const int c4 = threadIdx.x % cols4, rsub = threadIdx.x / cols4, rpb = 256 / cols4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long step = (long long)gridDim.x * rpb;
    for (long long r0 = (long long)blockIdx.x * rpb + rsub; r0 < rows; r0 += 4 * step) {  // four rows in flight
        float4 v[4];
        float sc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long r = r0 + k * step;
            const bool ok = r < rows;
            v[k] = ok ? *reinterpret_cast<const float4*>(src + r * ld_src + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            sc[k] = (ok && rowscale) ? __ldg(rowscale + r / rpg) : 1.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long r = r0 + k * step;
            if (r >= rows) break;
            float4 o = make_float4(v[k].x * sc[k], v[k].y * sc[k], v[k].z * sc[k], v[k].w * sc[k]);
            if (do_round) o = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
            reinterpret_cast<float4*>(dst)[r * cols4 + c4] = o;
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (rsub == 0) {
        for (int k = 1; k < rpb; ++k) {
            const float4 o = sh[k * cols4 + c4];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        reinterpret_cast<float4*>(partials)[(long long)blockIdx.x * cols4 + c4] = acc;
    }
}
// out[c] = sum_p partials[p][c]: 32 columns x 32 row slices per CTA, fixed order (deterministic)
__global__ void __launch_bounds__(1024) colsum_partials_kernel(const float* __restrict__ partials,
                                                               float* __restrict__ out, int P, int cols) {
    uwr_pdl_enter();
    __shared__ float sh[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < cols)
        for (int p = ty; p < P; p += 32) s += partials[(long long)p * cols + c];
    sh[ty][tx] = s;
    __syncthreads();
    const float v = warp_sum(sh[tx][ty]);
    const int co = blockIdx.x * 32 + ty;
    if (tx == 0 && co < cols) out[co] = v;
}

}  // namespace

extern "C" int uwr_scale_round_colsum(const float* src, long long ld_src, float* dst, long long rows, int cols,
                                      const float* rowscale, int rows_per_group, int do_round, float* colsum,
                                      float* workspace, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(src && dst && colsum && workspace && cols % 4 == 0 && ld_src % 4 == 0, "uwr_scale_round_colsum: bad args");
    UWR_REQUIRE(cols / 4 <= 256 && 256 % (cols / 4) == 0, "uwr_scale_round_colsum: cols/4 must divide 256");
    UWR_REQUIRE(!rowscale || rows_per_group > 0, "uwr_scale_round_colsum: rowscale needs rows_per_group");
    UWR_REQUIRE(rows > 0, "uwr_scale_round_colsum: empty");
    const int cols4 = cols / 4, rpb = 256 / cols4;
    long long blocks = (rows + rpb - 1) / rpb;
    long long cap = 8LL * uwr_sm_count();
    if (cap > 1024) cap = 1024;
    if (blocks > cap) blocks = cap;
    (void)uwr_launch_pdl(scale_round_colsum_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, src, ld_src, dst, rows, cols4, rowscale,
                                                                   rows_per_group > 0 ? rows_per_group : 1, do_round,
                                                                   workspace);
    UWR_CHECK_LAUNCH("scale_round_colsum_kernel");
    (void)uwr_launch_pdl(colsum_partials_kernel, dim3(uwr_cdiv(cols, 32)), dim3(1024), 0, stream, workspace, colsum, (int)blocks, cols);
    UWR_CHECK_LAUNCH("colsum_partials_kernel");
    return 0;
}

extern "C" int uwr_round_tf32_tensors(const float* const* src, float* const* dst, const long long* offsets,
                                      int n_tensors, long long total_elems, int do_round, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(src && dst && offsets && n_tensors > 0, "uwr_round_tf32_tensors: bad args");
    long long blocks = (total_elems + 4095) / 4096;
    if (blocks > 8LL * uwr_sm_count()) blocks = 8LL * uwr_sm_count();
    if (blocks < 1) blocks = 1;
    (void)uwr_launch_pdl(round_tensors_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, src, dst, offsets, n_tensors, total_elems, do_round);
    UWR_CHECK_LAUNCH("round_tensors_kernel");
    return 0;
}

extern "C" int uwr_scale_round(const float* src, long long ld_src, float* dst, long long rows, int cols,
                               const float* rowscale, int rows_per_group, int do_round, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(src && dst && cols % 4 == 0 && ld_src % 4 == 0, "uwr_scale_round: cols/ld must be multiples of 4");
    UWR_REQUIRE(!rowscale || rows_per_group > 0, "uwr_scale_round: rowscale needs rows_per_group");
    if (rows == 0) return 0;
    long long blocks = (rows * (cols / 4) + 255) / 256;
    if (blocks > 16LL * uwr_sm_count()) blocks = 16LL * uwr_sm_count();
    (void)uwr_launch_pdl(scale_round_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, src, ld_src, dst, rows, cols / 4, rowscale,
                                                            rows_per_group > 0 ? rows_per_group : 1, do_round);
    UWR_CHECK_LAUNCH("scale_round_kernel");
    return 0;
}
