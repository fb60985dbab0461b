// LayerNorm over the channel axis of token tensors (nn.LayerNorm, eps 1e-5, biased variance:
// AST.py:521,534,593,622).  One warp per row, rows kept in registers; HBM-bound:
// fwd 2*rows*C*4 B, bwd 3*rows*C*4 B (+ pass-through residual gradient).
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_PARTIAL_BLOCKS = 1024;

// VPL = values per lane (C = 32 * VPL when exact; general C handled by the bounds check)
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               float* __restrict__ y, float* __restrict__ mean,
                                                               float* __restrict__ rstd, long long rows, int C,
                                                               float eps, int rnd) {
    uwr_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float gm[VPL], bt[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + 32 * i;
        gm[i] = c < C ? gamma[c] : 0.f;
        bt[i] = c < C ? beta[c] : 0.f;
    }
    const float invC = 1.0f / (float)C;
    for (long long r = (long long)blockIdx.x * LN_WARPS + warp; r < rows; r += (long long)gridDim.x * LN_WARPS) {
        const float* xr = x + r * C;
        float v[VPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < C ? xr[c] : 0.f;
            s += v[i];
        }
        const float mu = warp_sum(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + 32 * i;
            const float d = c < C ? v[i] - mu : 0.f;
            q += d * d;
        }
        const float rs = rsqrtf(warp_sum(q) * invC + eps);
        float* yr = y + r * C;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float o = (v[i] - mu) * rs * gm[i] + bt[i];
                yr[c] = rnd ? tf32_round(o) : o;
            }
        }
        if (lane == 0) {
            if (mean) mean[r] = mu;
            if (rstd) rstd[r] = rs;
        }
    }
}

// Vectorised forward for C = 4 * LPR * V4: LPR lanes share a row (128-bit accesses, 32 / LPR rows per
// warp pass, RPI passes in flight), so a 128-byte row (C = 32) no longer costs a whole warp two
// full-width reductions.
template <int LPR, int V4>
__global__ void __launch_bounds__(LN_WARPS * 32, V4 <= 2 ? 4 : 2) ln_fwd_vec_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   float* __restrict__ y, float* __restrict__ mean,
                                                                   float* __restrict__ rstd, long long rows, float eps,
                                                                   int rnd) {
    uwr_pdl_enter();
    constexpr int C = 4 * LPR * V4;
    constexpr int RPW = 32 / LPR;               // rows per warp pass
    constexpr int RPI = V4 <= 2 ? 4 : (V4 <= 4 ? 2 : 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    float4 gm[V4], bt[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        gm[i] = *reinterpret_cast<const float4*>(gamma + (l + LPR * i) * 4);
        bt[i] = *reinterpret_cast<const float4*>(beta + (l + LPR * i) * 4);
    }
    constexpr float invC = 1.0f / (float)C;
    const long long stride = (long long)gridDim.x * LN_WARPS * RPW;
    // the loop bound is warp-uniform (first row of the warp's group): every lane takes part in the shuffles,
    // rows past the end are masked per lane
    for (long long rw = ((long long)blockIdx.x * LN_WARPS + warp) * RPW; rw < rows; rw += stride * RPI) {
        const long long r0 = rw + sub;
        float4 v[RPI][V4];
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
#pragma unroll
            for (int i = 0; i < V4; ++i)
                v[k][i] = r < rows ? *reinterpret_cast<const float4*>(x + r * C + (l + LPR * i) * 4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < V4; ++i) s += (v[k][i].x + v[k][i].y) + (v[k][i].z + v[k][i].w);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mu = s * invC;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < V4; ++i) {
                const float a = v[k][i].x - mu, b = v[k][i].y - mu, c = v[k][i].z - mu, d = v[k][i].w - mu;
                q += (a * a + b * b) + (c * c + d * d);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rs = rsqrtf(q * invC + eps);
            if (r < rows) {
#pragma unroll
                for (int i = 0; i < V4; ++i) {
                    float4 o4 = make_float4((v[k][i].x - mu) * rs * gm[i].x + bt[i].x, (v[k][i].y - mu) * rs * gm[i].y + bt[i].y,
                                            (v[k][i].z - mu) * rs * gm[i].z + bt[i].z, (v[k][i].w - mu) * rs * gm[i].w + bt[i].w);
                    if (rnd) o4 = make_float4(tf32_round(o4.x), tf32_round(o4.y), tf32_round(o4.z), tf32_round(o4.w));
                    *reinterpret_cast<float4*>(y + r * C + (l + LPR * i) * 4) = o4;
                }
                if (l == 0) {
                    if (mean) mean[r] = mu;
                    if (rstd) rstd[r] = rs;
                }
            }
        }
    }
}

template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const float* __restrict__ dy,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ dres,
                                                               float* __restrict__ dx, float* __restrict__ partials,
                                                               long long rows, int C) {
    uwr_pdl_enter();
    // RPI rows per warp iteration: all their loads are issued before the first reduction, which is
    // what keeps enough bytes in flight (one row per iteration ran at ~1.3 TB/s)
    constexpr int RPI = VPL <= 2 ? 4 : (VPL <= 8 ? 2 : 1);
    __shared__ float sh[LN_WARPS][32 * VPL];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float gm[VPL], dg[VPL], db[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = lane + 32 * i;
        gm[i] = c < C ? gamma[c] : 0.f;
        dg[i] = 0.f;
        db[i] = 0.f;
    }
    const float invC = 1.0f / (float)C;
    const long long stride = (long long)gridDim.x * LN_WARPS;
    for (long long r0 = (long long)blockIdx.x * LN_WARPS + warp; r0 < rows; r0 += stride * RPI) {
        float xv[RPI][VPL], dv[RPI][VPL], rv[RPI][VPL], mu[RPI], rs[RPI];
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
            const bool rok = r < rows;
            mu[k] = rok ? mean[r] : 0.f;
            rs[k] = rok ? rstd[r] : 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = lane + 32 * i;
                const bool ok = rok && c < C;
                xv[k][i] = ok ? x[r * C + c] : 0.f;
                dv[k][i] = ok ? dy[r * C + c] : 0.f;
                rv[k][i] = (ok && dres) ? dres[r * C + c] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
            if (r >= rows) break;
            float xh[VPL], g[VPL];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = lane + 32 * i;
                xh[i] = c < C ? (xv[k][i] - mu[k]) * rs[k] : 0.f;
                g[i] = dv[k][i] * gm[i];
                s1 += g[i];
                s2 += g[i] * xh[i];
                dg[i] += dv[k][i] * xh[i];
                db[i] += dv[k][i];
            }
            s1 = warp_sum(s1) * invC;
            s2 = warp_sum(s2) * invC;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = lane + 32 * i;
                if (c < C) dx[r * C + c] = rs[k] * (g[i] - s1 - xh[i] * s2) + rv[k][i];
            }
        }
    }
    // block-level reduction of dgamma/dbeta, then one partial row per block
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) sh[warp][lane + 32 * i] = pass == 0 ? dg[i] : db[i];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) s += sh[w][c];
            partials[((long long)blockIdx.x * 2 + pass) * C + c] = s;
        }
        __syncthreads();
    }
}

// Vectorised backward for C = 4 * LPR * V4 (same lane layout as ln_fwd_vec_kernel): 128-bit accesses, 32 / LPR
// rows per warp pass, two passes in flight for the narrow rows.  The warp-per-row kernel ran the 128-byte rows
// (C = 32) at 3.5 TB/s and the short, wide tensors of the deep stages (16 384 x 512) at 1.3 TB/s.
// DS: additionally emit d_s = tf32(rowscale[row / rpg] * dx) and its column sums — the DropPath-scaled, rounded copy
// of the residual-stream gradient that the NEXT backward function feeds to its GEMMs (it used to be a separate
// read + write pass, uwr_scale_round_colsum) — as a third partial row.
template <int LPR, int V4, bool DS>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_vec_kernel(const float* __restrict__ dy,
                                                                   const float* __restrict__ x,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd,
                                                                   const float* __restrict__ dres,
                                                                   float* __restrict__ dx, float* __restrict__ partials,
                                                                   long long rows, const float* __restrict__ ds_scale,
                                                                   int ds_rpg, float* __restrict__ ds_out, int ds_round) {
    uwr_pdl_enter();
    constexpr int C = 4 * LPR * V4;
    constexpr int RPW = 32 / LPR;
    constexpr int RPI = V4 == 1 ? 2 : 1;
    __shared__ __align__(16) float sh[LN_WARPS][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, l = lane % LPR;
    float4 gm[V4], dg[V4], db[V4], cs[DS ? V4 : 1];
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        gm[i] = *reinterpret_cast<const float4*>(gamma + (l + LPR * i) * 4);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (DS) cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr float invC = 1.0f / (float)C;
    const long long stride = (long long)gridDim.x * LN_WARPS * RPW;
    for (long long rw = ((long long)blockIdx.x * LN_WARPS + warp) * RPW; rw < rows; rw += stride * RPI) {
        const long long r0 = rw + sub;   // warp-uniform loop bound: every lane takes part in the shuffles
        float4 xv[RPI][V4], dv[RPI][V4], rv[RPI][V4];
        float mu[RPI], rs[RPI];
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
            const bool ok = r < rows;
            mu[k] = ok ? mean[r] : 0.f;
            rs[k] = ok ? rstd[r] : 0.f;
#pragma unroll
            for (int i = 0; i < V4; ++i) {
                const long long o = r * C + (l + LPR * i) * 4;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                xv[k][i] = ok ? *reinterpret_cast<const float4*>(x + o) : z;
                dv[k][i] = ok ? *reinterpret_cast<const float4*>(dy + o) : z;
                rv[k][i] = (ok && dres) ? *reinterpret_cast<const float4*>(dres + o) : z;
            }
        }
#pragma unroll
        for (int k = 0; k < RPI; ++k) {
            const long long r = r0 + k * stride;
            float4 xh[V4], g[V4];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < V4; ++i) {
                xh[i] = make_float4((xv[k][i].x - mu[k]) * rs[k], (xv[k][i].y - mu[k]) * rs[k],
                                    (xv[k][i].z - mu[k]) * rs[k], (xv[k][i].w - mu[k]) * rs[k]);
                g[i] = make_float4(dv[k][i].x * gm[i].x, dv[k][i].y * gm[i].y, dv[k][i].z * gm[i].z, dv[k][i].w * gm[i].w);
                s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
                dg[i].x += dv[k][i].x * xh[i].x; dg[i].y += dv[k][i].y * xh[i].y;
                dg[i].z += dv[k][i].z * xh[i].z; dg[i].w += dv[k][i].w * xh[i].w;
                db[i].x += dv[k][i].x; db[i].y += dv[k][i].y; db[i].z += dv[k][i].z; db[i].w += dv[k][i].w;
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            s1 *= invC;
            s2 *= invC;
            if (r < rows) {
                const float sc = (DS && ds_scale != nullptr) ? __ldg(ds_scale + r / ds_rpg) : 1.f;
#pragma unroll
                for (int i = 0; i < V4; ++i) {
                    const float4 o = make_float4(rs[k] * (g[i].x - s1 - xh[i].x * s2) + rv[k][i].x,
                                                 rs[k] * (g[i].y - s1 - xh[i].y * s2) + rv[k][i].y,
                                                 rs[k] * (g[i].z - s1 - xh[i].z * s2) + rv[k][i].z,
                                                 rs[k] * (g[i].w - s1 - xh[i].w * s2) + rv[k][i].w);
                    *reinterpret_cast<float4*>(dx + r * C + (l + LPR * i) * 4) = o;
                    if (DS) {
                        float4 q = make_float4(o.x * sc, o.y * sc, o.z * sc, o.w * sc);
                        if (ds_round) q = make_float4(tf32_round(q.x), tf32_round(q.y), tf32_round(q.z), tf32_round(q.w));
                        *reinterpret_cast<float4*>(ds_out + r * C + (l + LPR * i) * 4) = q;
                        cs[i].x += q.x; cs[i].y += q.y; cs[i].z += q.z; cs[i].w += q.w;
                    }
                }
            }
        }
    }
    // dgamma / dbeta [/ column sums of d_s]: lanes that share a column group (across the RPW sub-rows of the
    // warp), then the warps
    constexpr int NP = DS ? 3 : 2;
#pragma unroll
    for (int pass = 0; pass < NP; ++pass) {
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            float4 v = pass == 0 ? dg[i] : (pass == 1 ? db[i] : cs[DS ? i : 0]);
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
                v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
            }
            if (sub == 0) *reinterpret_cast<float4*>(&sh[warp][(l + LPR * i) * 4]) = v;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) t += sh[w][c];
            partials[((long long)blockIdx.x * NP + pass) * C + c] = t;
        }
        __syncthreads();
    }
}

// three-output variant of ln_param_reduce_kernel for the DS kernels (partial rows: dgamma, dbeta, colsum(d_s))
__global__ void __launch_bounds__(1024) ln_param_reduce3_kernel(const float* __restrict__ partials,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                float* __restrict__ colsum, int nblocks, int C) {
    uwr_pdl_enter();
    __shared__ float sh[3][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a[3] = {0.f, 0.f, 0.f};
    if (c < C)
        for (int i = ty; i < nblocks; i += 32) {
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] += partials[((long long)i * 3 + k) * C + c];
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) sh[k][ty][tx] = a[k];
    __syncthreads();
    const int co = blockIdx.x * 32 + ty;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = warp_sum(sh[k][tx][ty]);
        if (tx == 0 && co < C) (k == 0 ? dgamma : (k == 1 ? dbeta : colsum))[co] = v;
    }
}

// dgamma/dbeta = column sums of the per-CTA partial rows: 32 columns x 32 row-slices per CTA (a
// single thread per column walking ~900 partial rows serially took 70 us per LayerNorm)
__global__ void __launch_bounds__(1024) ln_param_reduce_kernel(const float* __restrict__ partials,
                                                               float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, int nblocks, int C) {
    uwr_pdl_enter();
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f, b = 0.f;
    if (c < C)
        for (int i = ty; i < nblocks; i += 32) {
            a += partials[((long long)i * 2 + 0) * C + c];
            b += partials[((long long)i * 2 + 1) * C + c];
        }
    sa[ty][tx] = a;
    sb[ty][tx] = b;
    __syncthreads();
    // warp ty reduces column ty (transposed read), fixed order => deterministic
    float va = sa[tx][ty], vb = sb[tx][ty];
    va = warp_sum(va);
    vb = warp_sum(vb);
    const int co = blockIdx.x * 32 + ty;
    if (tx == 0 && co < C) {
        dgamma[co] = va;
        dbeta[co] = vb;
    }
}

int ln_blocks(long long rows) {
    long long b = (rows + LN_WARPS - 1) / LN_WARPS;
    const long long cap = (long long)uwr_sm_count() * 6;
    if (b > cap) b = cap;
    if (b > LN_MAX_PARTIAL_BLOCKS) b = LN_MAX_PARTIAL_BLOCKS;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int uwr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                 float* rstd, long long rows, int C, float eps, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && gamma && beta && y, "uwr_layernorm_fwd: null pointer");
    UWR_REQUIRE(C > 0 && C <= 1024, "uwr_layernorm_fwd: C=%d unsupported (1..1024)", C);
    if (rows == 0) return 0;
    const int blocks = ln_blocks(rows);
#define LN_FWD_VEC(L, V)                                                                                         \
    do {                                                                                                         \
        long long b = (rows + LN_WARPS * (32 / L) - 1) / (LN_WARPS * (32 / L));                                  \
        const long long cap = (long long)uwr_sm_count() * (V <= 2 ? 4 : 2);   /* one resident wave */            \
        if (b > cap) b = cap;                                                                                    \
        (void)uwr_launch_pdl(ln_fwd_vec_kernel<L, V>, dim3((unsigned)b), dim3(LN_WARPS * 32), 0, stream, x, gamma, beta, y, mean, rstd, rows,  \
                                                                           eps, uwr_round_outputs());            \
        UWR_CHECK_LAUNCH("ln_fwd_vec_kernel");                                                                   \
        return 0;                                                                                                \
    } while (0)
    if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0) {
        switch (C) {
            case 16: LN_FWD_VEC(4, 1);
            case 32: LN_FWD_VEC(8, 1);
            case 64: LN_FWD_VEC(16, 1);
            case 128: LN_FWD_VEC(32, 1);
            case 256: LN_FWD_VEC(32, 2);
            case 512: LN_FWD_VEC(32, 4);
            case 1024: LN_FWD_VEC(32, 8);
            default: break;
        }
    }
#undef LN_FWD_VEC
    const int vpl = (C + 31) / 32;
#define LN_FWD(V) (void)uwr_launch_pdl(ln_fwd_kernel<V>, dim3(blocks), dim3(LN_WARPS * 32), 0, stream, x, gamma, beta, y, mean, rstd, rows, C, eps, uwr_round_outputs())
    if (vpl <= 1) LN_FWD(1);
    else if (vpl <= 2) LN_FWD(2);
    else if (vpl <= 4) LN_FWD(4);
    else if (vpl <= 8) LN_FWD(8);
    else if (vpl <= 16) LN_FWD(16);
    else LN_FWD(32);
#undef LN_FWD
    UWR_CHECK_LAUNCH("ln_fwd_kernel");
    return 0;
}

extern "C" size_t uwr_layernorm_bwd_workspace_bytes(long long rows, int C) {
    return (size_t)ln_blocks(rows) * 2 * (size_t)C * sizeof(float);
}

extern "C" int uwr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                                 float* partials, long long rows, int C, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && partials, "uwr_layernorm_bwd: null pointer");
    UWR_REQUIRE(C > 0 && C <= 1024, "uwr_layernorm_bwd: C=%d unsupported (1..1024)", C);
    UWR_REQUIRE(rows > 0, "uwr_layernorm_bwd: rows must be positive");
#define LN_BWD_VEC(L, V)                                                                                          \
    do {                                                                                                          \
        long long b = (rows + LN_WARPS * (32 / L) * 8 - 1) / (LN_WARPS * (32 / L) * 8); /* >= 8 rows per lane group */ \
        const long long cap = ln_blocks(rows);                                                                    \
        if (b > cap) b = cap;                                                                                     \
        if (b < 1) b = 1;                                                                                         \
        (void)uwr_launch_pdl(ln_bwd_vec_kernel<L, V, false>, dim3((unsigned)b), dim3(LN_WARPS * 32), 0, stream, dy, x, gamma, mean, rstd, dres, \
                                                                                  dx, partials, rows, nullptr, 1, \
                                                                                  nullptr, 0);                    \
        UWR_CHECK_LAUNCH("ln_bwd_vec_kernel");                                                                    \
        (void)uwr_launch_pdl(ln_param_reduce_kernel, dim3(uwr_cdiv(C, 32)), dim3(1024), 0, stream, partials, dgamma, dbeta, (int)b, C);         \
        UWR_CHECK_LAUNCH("ln_param_reduce_kernel");                                                               \
        return 0;                                                                                                 \
    } while (0)
    if ((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)(dres ? dres : x)) & 15) == 0) {
        switch (C) {
            case 16: LN_BWD_VEC(4, 1);
            case 32: LN_BWD_VEC(8, 1);
            case 64: LN_BWD_VEC(16, 1);
            case 128: LN_BWD_VEC(32, 1);
            case 256: LN_BWD_VEC(32, 2);
            case 512: LN_BWD_VEC(32, 4);
            default: break;
        }
    }
#undef LN_BWD_VEC
    const int blocks = ln_blocks(rows);
    const int vpl = (C + 31) / 32;
#define LN_BWD(V) \
    (void)uwr_launch_pdl(ln_bwd_kernel<V>, dim3(blocks), dim3(LN_WARPS * 32), 0, stream, dy, x, gamma, mean, rstd, dres, dx, partials, rows, C)
    if (vpl <= 1) LN_BWD(1);
    else if (vpl <= 2) LN_BWD(2);
    else if (vpl <= 4) LN_BWD(4);
    else if (vpl <= 8) LN_BWD(8);
    else if (vpl <= 16) LN_BWD(16);
    else LN_BWD(32);
#undef LN_BWD
    UWR_CHECK_LAUNCH("ln_bwd_kernel");
    (void)uwr_launch_pdl(ln_param_reduce_kernel, dim3(uwr_cdiv(C, 32)), dim3(1024), 0, stream, partials, dgamma, dbeta, blocks, C);
    UWR_CHECK_LAUNCH("ln_param_reduce_kernel");
    return 0;
}

// ---- LayerNorm backward that also emits the next backward function's GEMM operand ------------------------------
// d_s = tf32(rowscale[row / rows_per_group] * dx) (rounded only in single-pass TF32 mode) and ds_colsum = column
// sums of d_s (the bias gradient of the Linear that consumes d_s).  Served for C in {16, 32, 64, 128, 256, 512}.
extern "C" int uwr_layernorm_bwd_ds_supported(long long rows, int C) {
    return rows > 0 && (C == 16 || C == 32 || C == 64 || C == 128 || C == 256 || C == 512);
}

extern "C" size_t uwr_layernorm_bwd_ds_workspace_bytes(long long rows, int C) {
    return (size_t)ln_blocks(rows) * 3 * (size_t)C * sizeof(float);
}

extern "C" int uwr_layernorm_bwd_ds(const float* dy, const float* x, const float* gamma, const float* mean,
                                    const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                                    const float* ds_rowscale, int ds_rows_per_group, float* ds_out, float* ds_colsum,
                                    float* partials, long long rows, int C, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && partials && ds_out && ds_colsum,
                "uwr_layernorm_bwd_ds: null pointer");
    UWR_REQUIRE(uwr_layernorm_bwd_ds_supported(rows, C), "uwr_layernorm_bwd_ds: C=%d unsupported", C);
    UWR_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)ds_out |
                  (uintptr_t)(dres ? dres : x)) & 15) == 0, "uwr_layernorm_bwd_ds: pointers must be 16-byte aligned");
    UWR_REQUIRE(!ds_rowscale || ds_rows_per_group > 0, "uwr_layernorm_bwd_ds: rowscale needs rows_per_group");
    const int rpg = ds_rows_per_group > 0 ? ds_rows_per_group : 1;
#define LN_BWD_DS(L, V)                                                                                          \
    do {                                                                                                          \
        long long b = (rows + LN_WARPS * (32 / L) * 8 - 1) / (LN_WARPS * (32 / L) * 8);                           \
        const long long cap = ln_blocks(rows);                                                                    \
        if (b > cap) b = cap;                                                                                     \
        if (b < 1) b = 1;                                                                                         \
        (void)uwr_launch_pdl(ln_bwd_vec_kernel<L, V, true>, dim3((unsigned)b), dim3(LN_WARPS * 32), 0, stream,                                  \
            dy, x, gamma, mean, rstd, dres, dx, partials, rows, ds_rowscale, rpg, ds_out, uwr_round_outputs());   \
        UWR_CHECK_LAUNCH("ln_bwd_vec_kernel<DS>");                                                                \
        (void)uwr_launch_pdl(ln_param_reduce3_kernel, dim3(uwr_cdiv(C, 32)), dim3(1024), 0, stream, partials, dgamma, dbeta, ds_colsum, (int)b, C); \
        UWR_CHECK_LAUNCH("ln_param_reduce3_kernel");                                                              \
        return 0;                                                                                                 \
    } while (0)
    switch (C) {
        case 16: LN_BWD_DS(4, 1);
        case 32: LN_BWD_DS(8, 1);
        case 64: LN_BWD_DS(16, 1);
        case 128: LN_BWD_DS(32, 1);
        case 256: LN_BWD_DS(32, 2);
        default: LN_BWD_DS(32, 4);
    }
#undef LN_BWD_DS
}
