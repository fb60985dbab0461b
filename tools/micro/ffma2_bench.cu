// Packed fp32 (FFMA2 / FMUL2) issue-rate probe for sm_100a: the same 16 independent FMA chains per thread written
// as 16 scalar FFMA or 8 FFMA2, timed with CUDA events.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o ffma2_bench ffma2_bench.cu ; prints lane-FMA/clk/SM for both forms.
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) {
                acc[i] = __ffma2_rn(acc[i], a2, b2);
            } else {
                acc[i].x = fmaf(acc[i].x, a2.x, b2.x);
                acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, iters = 1 << 16;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (packed) probe<1><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
            else probe<0><<<blocks, 256>>>(out, iters, 0.999f, 1e-3f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)blocks * 256 * iters * 16;
            printf("%s rep %d: %.3f ms  %.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM at the %d MHz attribute clock)\n",
                   packed ? "FFMA2" : "FFMA ", rep, ms, fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
