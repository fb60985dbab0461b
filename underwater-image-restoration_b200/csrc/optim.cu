// Optimizer step of the training loop (src/ModelTrainer.py:87-88,197-204):
// clip_grad_norm_(params, 1.0) + Adam / AdamW over a table of tensors, two kernels per step
// instead of ~800 foreach launches.  HBM-bound: 7 * 4 B per parameter.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 4096;  // elements of the virtual concatenation per block iteration

__device__ __forceinline__ int find_tensor(const long long* __restrict__ offsets, int n, long long i) {
    int lo = 0, hi = n;  // offsets has n+1 entries, offsets[0] = 0
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(OPT_THREADS) grad_sqsum_kernel(const float* const* __restrict__ grads,
                                                                 const long long* __restrict__ offsets, int n,
                                                                 long long total, float prescale,
                                                                 float* __restrict__ partials) {
    uwr_pdl_enter();
    __shared__ float red[OPT_THREADS / 32];
    float s = 0.f;
    for (long long c0 = (long long)blockIdx.x * OPT_CHUNK; c0 < total; c0 += (long long)gridDim.x * OPT_CHUNK) {
        const long long c1 = min(total, c0 + OPT_CHUNK);
        long long i = c0 + threadIdx.x;
        if (i >= c1) continue;
        int t = find_tensor(offsets, n, i);
        for (; i < c1; i += OPT_THREADS) {
            while (i >= offsets[t + 1]) ++t;
            const float g = grads[t][i - offsets[t]] * prescale;
            s += g * g;
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) tsum += red[w];
        partials[blockIdx.x] = tsum;
    }
}

__global__ void grad_norm_final_kernel(const float* __restrict__ partials, int nblocks, float max_norm,
                                       float* __restrict__ out) {
    uwr_pdl_enter();
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) a += partials[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 32; ++i) s += sh[i];
        const float norm = (float)sqrt(s);
        out[0] = norm;
        const float coef = max_norm / (norm + 1e-6f);
        out[1] = coef < 1.f ? coef : 1.f;
    }
}

__global__ void __launch_bounds__(OPT_THREADS) adam_kernel(float* const* __restrict__ params,
                                                           const float* const* __restrict__ grads,
                                                           float* const* __restrict__ exp_avg,
                                                           float* const* __restrict__ exp_avg_sq,
                                                           const long long* __restrict__ offsets, int n,
                                                           long long total, const float* __restrict__ clip_coef,
                                                           float prescale, float lr, float beta1, float beta2,
                                                           float eps, float wd, int decoupled, int step,
                                                           const int* __restrict__ step_dev,
                                                           const float* __restrict__ lr_dev) {
    uwr_pdl_enter();
    const int st = step_dev ? *step_dev : step;
    if (lr_dev) lr = *lr_dev;   // device-resident learning rate: a captured CUDA graph follows scheduler updates
    const float gscale = prescale * (clip_coef ? *clip_coef : 1.f);
    const float bc1 = 1.f - powf(beta1, (float)st);
    const float bc2 = 1.f - powf(beta2, (float)st);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long c0 = (long long)blockIdx.x * OPT_CHUNK; c0 < total; c0 += (long long)gridDim.x * OPT_CHUNK) {
        const long long c1 = min(total, c0 + OPT_CHUNK);
        long long i = c0 + threadIdx.x;
        if (i >= c1) continue;
        int t = find_tensor(offsets, n, i);
        for (; i < c1; i += OPT_THREADS) {
            while (i >= offsets[t + 1]) ++t;
            const long long j = i - offsets[t];
            float p = params[t][j];
            float g = grads[t][j] * gscale;
            if (wd != 0.f) {
                if (decoupled) p *= (1.f - lr * wd);
                else g += wd * p;
            }
            const float m = beta1 * exp_avg[t][j] + (1.f - beta1) * g;
            const float v = beta2 * exp_avg_sq[t][j] + (1.f - beta2) * g * g;
            exp_avg[t][j] = m;
            exp_avg_sq[t][j] = v;
            const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
            params[t][j] = p - step_size * (m / denom);
        }
    }
}

__global__ void increment_kernel(int* p) {
    uwr_pdl_enter(); *p += 1; }

int opt_blocks(long long total) {
    long long b = (total + OPT_CHUNK - 1) / OPT_CHUNK;
    const long long cap = 8LL * uwr_sm_count();
    if (b > cap) b = cap;
    if (b > 4096) b = 4096;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int uwr_grad_norm(const float* const* grads, const long long* offsets, int n_tensors, long long total_elems,
                             float max_norm, float grad_prescale, float* norm_out, float* workspace,
                             uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(grads && offsets && norm_out && workspace && n_tensors > 0, "uwr_grad_norm: bad args");
    const int blocks = opt_blocks(total_elems);
    (void)uwr_launch_pdl(grad_sqsum_kernel, dim3(blocks), dim3(OPT_THREADS), 0, stream, grads, offsets, n_tensors, total_elems, grad_prescale, workspace);
    UWR_CHECK_LAUNCH("grad_sqsum_kernel");
    (void)uwr_launch_pdl(grad_norm_final_kernel, dim3(1), dim3(32), 0, stream, workspace, blocks, max_norm, norm_out);
    UWR_CHECK_LAUNCH("grad_norm_final_kernel");
    return 0;
}

extern "C" int uwr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const long long* offsets, int n_tensors, long long total_elems,
                             const float* clip_coef, float grad_prescale, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int decoupled, int step, const int* step_dev, const float* lr_dev,
                             uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(params && grads && exp_avg && exp_avg_sq && offsets && n_tensors > 0, "uwr_adam_step: bad args");
    UWR_REQUIRE(step_dev || step >= 1, "uwr_adam_step: step must be >= 1");
    (void)uwr_launch_pdl(adam_kernel, dim3(opt_blocks(total_elems)), dim3(OPT_THREADS), 0, stream, params, grads, exp_avg, exp_avg_sq, offsets,
                                                                    n_tensors, total_elems, clip_coef, grad_prescale,
                                                                    lr, beta1, beta2, eps, weight_decay, decoupled,
                                                                    step, step_dev, lr_dev);
    UWR_CHECK_LAUNCH("adam_kernel");
    return 0;
}

extern "C" int uwr_increment_i32(int* counter, uwr_stream_t stream_) {
    (void)uwr_launch_pdl(increment_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream_, counter);
    UWR_CHECK_LAUNCH("increment_kernel");
    return 0;
}
