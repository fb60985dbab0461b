// Shared device helpers for the uwr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// Error plumbing (uwr_abi.cu): every C-ABI entry returns 0 or a negative code and leaves a
// message retrievable through uwr_last_error().
extern "C" const char* uwr_last_error(void);
void uwr_set_error(const char* fmt, ...);

#define UWR_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            uwr_set_error(__VA_ARGS__);        \
            return -1;                         \
        }                                      \
    } while (0)

extern int g_uwr_gemm_passes;
// producers round GEMM-operand outputs to TF32 at their stores in single-pass mode
static inline int uwr_round_outputs() { return g_uwr_gemm_passes == 1; }
extern unsigned long long g_uwr_launches;  // kernels launched by this library (bench evidence)

#define UWR_CHECK_LAUNCH(name)                                               \
    do {                                                                     \
        ++g_uwr_launches;                                                    \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) {                                            \
            uwr_set_error("%s: %s", name, cudaGetErrorString(e__));          \
            return -2;                                                       \
        }                                                                    \
    } while (0)

#define UWR_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) {                                             \
            uwr_set_error("%s: %s", #call, cudaGetErrorString(e__));          \
            return -2;                                                        \
        }                                                                     \
    } while (0)

static inline int uwr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int uwr_sm_count();

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// A training step is ~600 dependent kernels on one stream; between two of them the GPU drains, the next grid is
// launched, its CTAs run their set-up (barrier init, TMEM allocation, weight / table loads into registers).  With PDL the
// next grid is launched as soon as every CTA of the running one has started (griddepcontrol.launch_dependents, the first
// instruction of every converted kernel), its CTAs become resident wherever an SM has room -- first in the tail of the
// running kernel -- do whatever needs no global data, and block in griddepcontrol.wait until the WHOLE preceding grid has
// completed and its memory is visible.  Rules that keep this equivalent to plain stream order:
//   * a kernel launched through uwr_launch_pdl() executes uwr_pdl_wait() (or uwr_pdl_enter()) on every path before its
//     first global-memory access of any kind -- reads of a predecessor's output AND writes a predecessor might still read;
//     kernel parameters (tensor maps included) are not global memory in this sense;
//   * everything else keeps the <<<>>> launch: it then starts after the converted kernel before it has completed, and a
//     converted kernel after it is launched when it completes (no trigger = trigger at exit).
// Stream capture turns the attribute into a programmatic edge of the CUDA graph.  Without the attribute (the default:
// UWR_PDL unset / uwr_set_pdl(0)) the same launches are plain stream-ordered launches and the device instructions no-ops.
__device__ __forceinline__ void uwr_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void uwr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void uwr_pdl_enter() {
    uwr_pdl_trigger();
    uwr_pdl_wait();
}
int uwr_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t uwr_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = uwr_pdl_enabled();
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// ---- numerics ---------------------------------------------------------------------------
// fp32 -> tf32 with round-to-nearest (ties away); the tensor core would otherwise truncate,
// which biases every product by ~-1e-3 relative.
// cvt.rna.tf32.f32 is not a native sm_100a instruction: ptxas expands it to ~6 instructions
// (FSETP / IADD / LOP3 / SEL ... for the NaN and overflow cases) and it was 40 % of the attention
// backward's instruction stream.  On the sign-magnitude bit pattern "add half an ulp of the 10-bit
// mantissa, clear the low 13 bits" is the same rounding (nearest, ties away from zero) for every
// finite value below the overflow threshold, in two integer instructions.
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float(f2tf32(x)); }

// erf-form GELU as nn.GELU() in the reference (AST.py:295-301).  Phi(x) and phi(x) share ONE
// exp(-x^2/2): erf(z), z = |x|/sqrt(2), by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32
// rounding level) needs exp(-z^2) = exp(-x^2/2), which is also the Gaussian density.  ~14
// instructions for gelu AND gelu' instead of two erff + one expf (the dwconv kernels were
// ALU-bound on erff).
// MUFU without the denormal-range fix-up code that __expf / __fdividef expand to without -ftz (the
// fix-ups were ~40 % of the instructions of the issue-bound dwconv kernels): arguments here are
// never denormal-sensitive (1 <= 1 + p|x|; exp(-x^2/2) flushing to 0 below 1e-38 is exact enough).
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
    const float t = rcp_ftz(fmaf(0.23164189f, fabsf(x), 1.0f));  // 1/(1 + p*z), p*z = 0.3275911*|x|/sqrt(2)
    const float s = x * 0.84932180028801904272f;                 // sqrt(log2(e)/2): exp(-x^2/2) = 2^(-s^2)
    const float e = ex2_ftz(-(s * s));
    float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);  // the 0.5 of Phi(-|x|) folded in
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, 0.5f * -0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    const float tail = (poly * t) * e;  // = 0.5*(1 - erf(z)) = Phi(-|x|)
    cdf = x >= 0.f ? 1.0f - tail : tail;
    pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_f(float x) {
    float cdf, pdf;
    gelu_parts(x, cdf, pdf);
    return x * cdf;
}
// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): per-lane results identical to the scalar IEEE forms, one
// issue slot for two lanes (tools/micro/ffma2_bench.cu: 113 vs 73 lane-FMA/clk/SM on independent chains).  Used where
// it measured faster on the same box: the dwconv backward (-27 % instructions, 0.619 -> 0.590 ms at B16 H256 Ch256);
// the dwconv forward got slower with it (0.523 -> 0.560 ms) and stays scalar.
#ifndef UWR_PACKED_FP32
#define UWR_PACKED_FP32 1   // 0: the same expressions as scalar instructions (A/B builds only)
#endif
#if UWR_PACKED_FP32
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
#else
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif
__device__ __forceinline__ float2 splat2(float s) { return make_float2(s, s); }
// gelu_parts on a pair: the two MUFUs per lane, |x| and the sign transfer stay scalar; the polynomial and the products
// are packed (19 issue slots per pair instead of ~32).  cdf = 1/2 + copysign(1/2 - Phi(-|x|), x): absolute error
// <= 2^-25 against the select form, far below the 1.5e-7 of the erf approximation itself.
__device__ __forceinline__ void gelu_parts2(float2 x, float2& cdf, float2& pdf) {
    float2 t, e;
    t.x = rcp_ftz(fmaf(0.23164189f, fabsf(x.x), 1.0f));
    t.y = rcp_ftz(fmaf(0.23164189f, fabsf(x.y), 1.0f));
    const float2 a = fmul2(fmul2(x, x), splat2(-0.72134752044448170368f));   // -x^2 log2(e) / 2
    e.x = ex2_ftz(a.x);
    e.y = ex2_ftz(a.y);
    float2 poly = ffma2(splat2(0.5f * 1.061405429f), t, splat2(0.5f * -1.453152027f));
    poly = ffma2(poly, t, splat2(0.5f * 1.421413741f));
    poly = ffma2(poly, t, splat2(0.5f * -0.284496736f));
    poly = ffma2(poly, t, splat2(0.5f * 0.254829592f));
    const float2 tail = fmul2(fmul2(poly, t), e);                            // Phi(-|x|)
    const float2 h = ffma2(tail, splat2(-1.0f), splat2(0.5f));               // 1/2 - Phi(-|x|) >= 0
    float2 hs;
    hs.x = __uint_as_float(__float_as_uint(h.x) | (__float_as_uint(x.x) & 0x80000000u));
    hs.y = __uint_as_float(__float_as_uint(h.y) | (__float_as_uint(x.y) & 0x80000000u));
    cdf = fadd2(hs, splat2(0.5f));
    pdf = fmul2(e, splat2(0.39894228040143267794f));
}
__device__ __forceinline__ float2 gelu_f2(float2 x) {
    float2 cdf, pdf;
    gelu_parts2(x, cdf, pdf);
    return fmul2(x, cdf);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    float cdf, pdf;
    gelu_parts(x, cdf, pdf);
    return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// m16n8k8 TF32 tensor-core MMA, fp32 accumulate (legacy warp-level path; used where the tile
// shape is far below a tcgen05 atom, e.g. 64x64x32 attention windows).
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
    // not volatile: a pure function of its operands, so the scheduler may interleave independent MMAs
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
