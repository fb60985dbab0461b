// Cost of a dependent kernel node inside a replayed CUDA graph, with and without programmatic dependent launch (PDL):
// a chain of N launches captured from a stream, (a) plain, (b) every launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization and the kernel does griddepcontrol.launch_dependents + .wait first.
// Two kernel sizes: one tiny CTA (pure launch gap) and a 148-CTA persistent-style kernel with a setup phase (barrier
// init / shared-memory fill, what the GEMM does before it touches global memory) in front of a short body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_gap pdl_gap.cu && ./pdl_gap
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) body(const float* __restrict__ in, float* __restrict__ out, int n, int setup) {
    extern __shared__ float sh[];
    if (PDL) asm volatile("griddepcontrol.launch_dependents;");
    for (int i = threadIdx.x; i < setup; i += blockDim.x) sh[i] = (float)i;   // setup that needs no global data
    __syncthreads();
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] + sh[i % (setup > 0 ? setup : 1)];
}

template <bool PDL>
float run(int N, int grid, int n, int setup, float* a, float* b) {
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaFuncSetAttribute(body<PDL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 64 * 1024; cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = PDL ? 1 : 0;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        const float* in = (i & 1) ? b : a;
        float* out = (i & 1) ? a : b;
        cudaError_t e = cudaLaunchKernelEx(&cfg, body<PDL>, in, out, n, setup);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -1.f; }
    }
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (e != cudaSuccess) { printf("capture: %s\n", cudaGetErrorString(e)); return -1.f; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return -1.f; }
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
    return ms / 5 / N * 1000.f;
}

int main() {
    const int n = 1 << 22;
    float *a, *b;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
    cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
    struct { const char* name; int grid, n, setup; } cases[] = {
        {"1 CTA, 1 K elements, no setup", 1, 1024, 0},
        {"148 CTAs, 64 K elements, 16 K-float setup", 148, 65536, 16384},
        {"148 CTAs, 4 M elements, 16 K-float setup", 148, n, 16384},
        {"592 CTAs, 4 M elements, no setup", 592, n, 0},
    };
    for (auto& c : cases) {
        const float plain = run<false>(400, c.grid, c.n, c.setup, a, b);
        const float pdl = run<true>(400, c.grid, c.n, c.setup, a, b);
        printf("%-44s plain %6.2f us per node   PDL %6.2f us per node\n", c.name, plain, pdl);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
