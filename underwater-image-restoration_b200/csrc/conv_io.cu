// Convolutions at the model boundary and between scales, all on token (NHWC) tensors:
//   InputProj  Conv3x3(3->C)+LeakyReLU   (AST.py:447-466)      direct, lane <-> output channel
//   OutputProj Conv3x3(C->3) + residual  (AST.py:470-493,921)  direct, lanes split input channels
//   Downsample Conv4x4 s2 p1             (AST.py:408-424)      im2col (ky,kx,ci) + uwr_gemm_tf32
//   Upsample   ConvTranspose2x2 s2       (AST.py:428-443)      uwr_gemm_tf32 + 2x2 pixel scatter
// plus the strided copy used for the skip concatenation (AST.py:904-916) and a deterministic
// column sum (bias gradients).  All of these are HBM-bound data movement around the GEMMs.
#include <cstdlib>
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// InputProj
constexpr int IP_TS = 16;
constexpr int IP_HS = IP_TS + 2;
constexpr int IP_MAXCIN = 4;

template <int NCO>  // output channels per lane (Cout = 32*NCO)
__global__ void __launch_bounds__(256) input_proj_fwd_kernel(const float* __restrict__ img,
                                                             const float* __restrict__ weight,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ tokens, int H, int W, int Cin,
                                                             int Cout, float slope, int tiles_x) {
    uwr_pdl_enter();
    __shared__ float in_s[IP_MAXCIN][IP_HS][IP_HS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * IP_TS, tx0 = (blockIdx.x % tiles_x) * IP_TS;
    for (int idx = threadIdx.x; idx < Cin * IP_HS * IP_HS; idx += 256) {
        const int ci = idx / (IP_HS * IP_HS), rem = idx % (IP_HS * IP_HS);
        const int y = ty0 + rem / IP_HS - 1, x = tx0 + rem % IP_HS - 1;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = img[(((long long)b * Cin + ci) * H + y) * W + x];
        in_s[ci][rem / IP_HS][rem % IP_HS] = v;
    }
    float w[NCO][IP_MAXCIN * 9], bv[NCO];
#pragma unroll
    for (int n = 0; n < NCO; ++n) {
        const int co = lane + 32 * n;
        bv[n] = bias[co];
#pragma unroll
        for (int k = 0; k < IP_MAXCIN * 9; ++k) w[n][k] = k < Cin * 9 ? weight[co * Cin * 9 + k] : 0.f;
    }
    __syncthreads();
    for (int pp = 0; pp < 32; ++pp) {
        const int pix = warp * 32 + pp;
        const int ly = pix / IP_TS, lx = pix % IP_TS;
        const int y = ty0 + ly, x = tx0 + lx;
        if (y >= H || x >= W) continue;
        float acc[NCO];
#pragma unroll
        for (int n = 0; n < NCO; ++n) acc[n] = bv[n];
#pragma unroll
        for (int ci = 0; ci < IP_MAXCIN; ++ci) {
            if (ci >= Cin) break;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float iv = in_s[ci][ly + ky][lx + kx];
#pragma unroll
                    for (int n = 0; n < NCO; ++n) acc[n] = fmaf(iv, w[n][ci * 9 + ky * 3 + kx], acc[n]);
                }
        }
        const long long tok = ((long long)b * H + y) * W + x;
#pragma unroll
        for (int n = 0; n < NCO; ++n) {
            const float a = acc[n];
            tokens[tok * Cout + lane + 32 * n] = a > 0.f ? a : a * slope;
        }
    }
}

template <int NCO>
__global__ void __launch_bounds__(256) input_proj_bwd_kernel(const float* __restrict__ dtokens,
                                                             const float* __restrict__ tokens,
                                                             const float* __restrict__ img,
                                                             float* __restrict__ partials, int B, int H, int W,
                                                             int Cin, int Cout, float slope, int tiles_x,
                                                             int tiles_per_img) {
    uwr_pdl_enter();
    __shared__ float in_s[IP_MAXCIN][IP_HS][IP_HS];
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dw[NCO][IP_MAXCIN * 9 + 1];
#pragma unroll
    for (int n = 0; n < NCO; ++n)
#pragma unroll
        for (int k = 0; k < IP_MAXCIN * 9 + 1; ++k) dw[n][k] = 0.f;

    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * IP_TS, tx0 = (tl % tiles_x) * IP_TS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < Cin * IP_HS * IP_HS; idx += 256) {
            const int ci = idx / (IP_HS * IP_HS), rem = idx % (IP_HS * IP_HS);
            const int y = ty0 + rem / IP_HS - 1, x = tx0 + rem % IP_HS - 1;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = img[(((long long)b * Cin + ci) * H + y) * W + x];
            in_s[ci][rem / IP_HS][rem % IP_HS] = v;
        }
        __syncthreads();
        for (int pp = 0; pp < 32; ++pp) {
            const int pix = warp * 32 + pp;
            const int ly = pix / IP_TS, lx = pix % IP_TS;
            const int y = ty0 + ly, x = tx0 + lx;
            if (y >= H || x >= W) continue;
            const long long tok = ((long long)b * H + y) * W + x;
            float d[NCO];
#pragma unroll
            for (int n = 0; n < NCO; ++n) {
                const float o = tokens[tok * Cout + lane + 32 * n];
                d[n] = dtokens[tok * Cout + lane + 32 * n] * (o > 0.f ? 1.f : slope);
                dw[n][IP_MAXCIN * 9] += d[n];
            }
#pragma unroll
            for (int ci = 0; ci < IP_MAXCIN; ++ci) {
                if (ci >= Cin) break;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float iv = in_s[ci][ly + ky][lx + kx];
#pragma unroll
                        for (int n = 0; n < NCO; ++n) dw[n][ci * 9 + ky * 3 + kx] = fmaf(iv, d[n], dw[n][ci * 9 + ky * 3 + kx]);
                    }
            }
        }
    }
    // cross-warp reduction, one k at a time; partial layout [cta][k][Cout], k = Cin*9 -> bias
    const int nk = Cin * 9 + 1;
#pragma unroll
    for (int n = 0; n < NCO; ++n)
#pragma unroll
        for (int k = 0; k < IP_MAXCIN * 9 + 1; ++k) {
            const int kk = (k == IP_MAXCIN * 9) ? Cin * 9 : k;
            if (k != IP_MAXCIN * 9 && k >= Cin * 9) continue;
            __syncthreads();
            red[warp][lane] = dw[n][k];
            __syncthreads();
            if (warp == 0) {
                float s = 0.f;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) s += red[ww][lane];
                partials[((long long)blockIdx.x * nk + kk) * Cout + lane + 32 * n] = s;
            }
        }
}

__global__ void input_proj_reduce_kernel(const float* __restrict__ partials, float* __restrict__ dweight,
                                         float* __restrict__ dbias, int P, int Cin, int Cout) {
    uwr_pdl_enter();
    const int nk = Cin * 9 + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nk * Cout) return;
    const int k = idx / Cout, co = idx % Cout;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += partials[((long long)p * nk + k) * Cout + co];
    if (k < Cin * 9) dweight[co * Cin * 9 + k] = s;
    else dbias[co] = s;
}

// ------------------------------------------------------------------------------------------
// OutputProj
constexpr int OP_TY = 8, OP_TX = 16;
constexpr int OP_HY = OP_TY + 2, OP_HX = OP_TX + 2;

template <int CPL>  // input channels per lane (Cin = 32*CPL)
__global__ void __launch_bounds__(256) output_proj_fwd_kernel(const float* __restrict__ tokens, long long ld,
                                                              const float* __restrict__ weight,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ residual,
                                                              float* __restrict__ out, int H, int W, int tiles_x) {
    uwr_pdl_enter();
    constexpr int Cin = 32 * CPL;
    extern __shared__ __align__(16) float smem[];  // [OP_HY*OP_HX][Cin]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int ty0 = (blockIdx.x / tiles_x) * OP_TY, tx0 = (blockIdx.x % tiles_x) * OP_TX;
    for (int idx = threadIdx.x; idx < OP_HY * OP_HX * (Cin / 4); idx += 256) {
        const int pix = idx / (Cin / 4), c4 = (idx % (Cin / 4)) * 4;
        const int y = ty0 + pix / OP_HX - 1, x = tx0 + pix % OP_HX - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W)
            v = *reinterpret_cast<const float4*>(tokens + (((long long)b * H + y) * W + x) * ld + c4);
        *reinterpret_cast<float4*>(smem + pix * Cin + c4) = v;
    }
    float w[3][CPL][9];
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int e = 0; e < CPL; ++e)
#pragma unroll
            for (int k = 0; k < 9; ++k) w[co][e][k] = weight[(co * Cin + lane * CPL + e) * 9 + k];
    __syncthreads();
    const int ly = warp;  // one tile row per warp
    const int y = ty0 + ly;
    float keep[3] = {0.f, 0.f, 0.f};
    for (int lx = 0; lx < OP_TX; ++lx) {
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* sp = smem + ((ly + ky) * OP_HX + lx + kx) * Cin + lane * CPL;
#pragma unroll
                for (int e = 0; e < CPL; ++e) {
                    const float xv = sp[e];
#pragma unroll
                    for (int co = 0; co < 3; ++co) acc[co] = fmaf(xv, w[co][e][ky * 3 + kx], acc[co]);
                }
            }
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            const float s = warp_sum(acc[co]);
            if (lane == lx) keep[co] = s;
        }
    }
    const int x = tx0 + lane;
    if (lane < OP_TX && y < H && x < W) {
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            const long long o = (((long long)b * 3 + co) * H + y) * W + x;
            float v = keep[co] + bias[co];
            if (residual) v += residual[o];
            out[o] = v;
        }
    }
}

template <int CPL>
__global__ void __launch_bounds__(256) output_proj_bwd_kernel(const float* __restrict__ dout,
                                                              const float* __restrict__ tokens, long long ld,
                                                              const float* __restrict__ weight,
                                                              float* __restrict__ dtokens,
                                                              float* __restrict__ partials, int B, int H, int W,
                                                              int tiles_x, int tiles_per_img) {
    uwr_pdl_enter();
    constexpr int Cin = 32 * CPL;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                           // [OP_HY*OP_HX][Cin]
    float* dys = smem + OP_HY * OP_HX * Cin;    // [3][OP_HY][OP_HX]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w[3][CPL][9], dw[3][CPL][9], db[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int e = 0; e < CPL; ++e)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                w[co][e][k] = weight[(co * Cin + lane * CPL + e) * 9 + k];
                dw[co][e][k] = 0.f;
            }
    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * OP_TY, tx0 = (tl % tiles_x) * OP_TX;
        __syncthreads();
        for (int idx = threadIdx.x; idx < OP_HY * OP_HX * (Cin / 4); idx += 256) {
            const int pix = idx / (Cin / 4), c4 = (idx % (Cin / 4)) * 4;
            const int y = ty0 + pix / OP_HX - 1, x = tx0 + pix % OP_HX - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H && x >= 0 && x < W)
                v = *reinterpret_cast<const float4*>(tokens + (((long long)b * H + y) * W + x) * ld + c4);
            *reinterpret_cast<float4*>(xs + pix * Cin + c4) = v;
        }
        for (int idx = threadIdx.x; idx < 3 * OP_HY * OP_HX; idx += 256) {
            const int co = idx / (OP_HY * OP_HX), rem = idx % (OP_HY * OP_HX);
            const int y = ty0 + rem / OP_HX - 1, x = tx0 + rem % OP_HX - 1;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = dout[(((long long)b * 3 + co) * H + y) * W + x];
            dys[idx] = v;
        }
        __syncthreads();
        const int ly = warp;
        const int y = ty0 + ly;
        if (y < H) {
            for (int lx = 0; lx < OP_TX; ++lx) {
                const int x = tx0 + lx;
                if (x >= W) break;
                float dx[CPL];
#pragma unroll
                for (int e = 0; e < CPL; ++e) dx[e] = 0.f;
                float dyc[3];
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    dyc[co] = dys[(co * OP_HY + ly + 1) * OP_HX + lx + 1];
                    db[co] += dyc[co];
                }
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* sp = xs + ((ly + ky) * OP_HX + lx + kx) * Cin + lane * CPL;
#pragma unroll
                        for (int co = 0; co < 3; ++co) {
                            // dX[y,x] += dY[co][y+1-ky][x+1-kx] * W[co][ci][ky][kx]
                            const float dyv = dys[(co * OP_HY + ly + 2 - ky) * OP_HX + lx + 2 - kx];
#pragma unroll
                            for (int e = 0; e < CPL; ++e) {
                                dx[e] = fmaf(dyv, w[co][e][ky * 3 + kx], dx[e]);
                                dw[co][e][ky * 3 + kx] = fmaf(dyc[co], sp[e], dw[co][e][ky * 3 + kx]);
                            }
                        }
                    }
                float* dp = dtokens + (((long long)b * H + y) * W + x) * Cin + lane * CPL;
#pragma unroll
                for (int e = 0; e < CPL; ++e) dp[e] = dx[e];
            }
        }
    }
    // cross-warp reduction through shared memory: partial layout [cta][27*Cin + 3]
    __syncthreads();
    float* red = smem;  // [8][27*Cin]  (27*64*8*4 = 55 KB <= tile bytes? no -> reduce per co)
    for (int co = 0; co < 3; ++co) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < CPL; ++e)
#pragma unroll
            for (int k = 0; k < 9; ++k) red[(warp * Cin + lane * CPL + e) * 9 + k] = dw[co][e][k];
        __syncthreads();
        for (int idx = threadIdx.x; idx < Cin * 9; idx += 256) {
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += red[ww * Cin * 9 + idx];
            partials[(long long)blockIdx.x * (27 * Cin + 3) + co * Cin * 9 + idx] = s;
        }
    }
    __syncthreads();
    if (lane == 0) red[warp * 3 + 0] = db[0], red[warp * 3 + 1] = db[1], red[warp * 3 + 2] = db[2];
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int ww = 0; ww < 8; ++ww) s += red[ww * 3 + threadIdx.x];
        partials[(long long)blockIdx.x * (27 * Cin + 3) + 27 * Cin + threadIdx.x] = s;
    }
}

// OutputProj backward on the tensor cores (single-pass TF32 mode).  With the im2col of the 3-channel cotangent
//   A[p][k] = dY[co][py + 1 - ky][px + 1 - kx],   k = co*9 + ky*3 + kx  (27 columns, padded to 32)
// both gradients are thin GEMMs over the pixels p of a tile that share A:
//   dX[p][ci]  = sum_k A[p][k] Wm[k][ci]            (M = 16 pixels per warp, N = Cin, K = 32; 3xTF32: fp32-level)
//   dW[k][ci] += sum_p A[p][k] X[p][ci]             (M = 32, N = 8 channels per warp, K = the tile's pixels; 3xTF32)
// A is never materialised: its fragments are read straight from the 3 x 10 x 18 halo tile of dY.  The scalar kernel
// above spends 54 FMAs per (pixel, channel) on the CUDA cores (0.70 ms for B=16 256x256 Cin=64, 8x its HBM time);
// here the CUDA cores only stage and round the operands.
constexpr int OPM_XS = 8;
// 3xTF32 operand split: hi = the value rounded to TF32, lo = the exact remainder (the tensor core truncates it again: < 2^-21)
__device__ __forceinline__ void split_hi_lo(float v, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(v);
    lo = __float_as_uint(v - __uint_as_float(hi));
}   // row padding of the X / Wm tiles: fragment loads (k = t, n = g) hit 32 distinct banks

template <int CIN>
__global__ void __launch_bounds__(256, 3) output_proj_bwd_mma_kernel(const float* __restrict__ dout,
                                                                     const float* __restrict__ tokens, long long ld,
                                                                     const float* __restrict__ weight,
                                                                     float* __restrict__ dtokens,
                                                                     float* __restrict__ partials, int B, int H, int W,
                                                                     int tiles_x, int tiles_per_img) {
    uwr_pdl_enter();
    constexpr int XS = CIN + OPM_XS;
    constexpr int NT = CIN / 8;          // 8-channel column tiles
    constexpr int KSPLIT = 8 / NT;       // warps sharing a column tile split the tile's pixels (Cin = 32: two halves)
    constexpr int NPIX = OP_TY * OP_TX;  // 128
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                     // [128][XS]   X at the tile's pixels, full fp32
    float* wm = xs + NPIX * XS;           // [32][XS]    Wm[k][ci], TF32-rounded (hi part), rows 27..31 zero
    float* wl = wm + 32 * XS;             // [32][XS]    Wm - hi: the data gradient is 3xTF32 (every gradient of the network
                                          //             flows through it; single-pass TF32 here doubled the model's gradient error)
    float* dyf = wl + 32 * XS;            // [3][10][18] dY halo tile in full fp32, zero outside the image (hi / lo split at the
                                          //             fragment loads: both gradients are 3xTF32)
    __shared__ float dbred[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;

    for (int idx = threadIdx.x; idx < 32 * CIN; idx += 256) {
        const int k = idx / CIN, ci = idx % CIN;
        float v = 0.f;
        if (k < 27) v = weight[((k / 9) * CIN + ci) * 9 + k % 9];
        const float hi = tf32_round(v);
        wm[k * XS + ci] = hi;
        wl[k * XS + ci] = v - hi;
    }
    // offsets of this lane's A columns inside the halo tile: k = ks*8 + t (+4); columns >= 27 are zero
    int aoff[8];
    unsigned avalid = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = (j >> 1) * 8 + t + (j & 1) * 4;
        const int kk = k < 27 ? k : 0, co = kk / 9, ky = (kk % 9) / 3, kx = kk % 3;
        aoff[j] = co * (OP_HY * OP_HX) + (2 - ky) * OP_HX + (2 - kx);
        if (k < 27) avalid |= 1u << j;
    }
    // ... and of its A^T rows m = mt*16 + g (+8) for the weight gradient
    int moff[4];
    unsigned mvalid = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = (j >> 1) * 16 + g + (j & 1) * 8;
        const int mm = m < 27 ? m : 0, co = mm / 9, ky = (mm % 9) / 3, kx = mm % 3;
        moff[j] = co * (OP_HY * OP_HX) + (2 - ky) * OP_HX + (2 - kx);
        if (m < 27) mvalid |= 1u << j;
    }
    const int nt2 = warp % NT, kh = warp / NT;   // weight-gradient role of this warp
    float dwacc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int c = 0; c < 4; ++c) dwacc[mt][c] = 0.f;
    float db0 = 0.f, db1 = 0.f, db2 = 0.f;

    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * OP_TY, tx0 = (tl % tiles_x) * OP_TX;
        __syncthreads();   // the previous tile's readers are done (also orders the wm fill before its first use)
        for (int idx = threadIdx.x; idx < NPIX * (CIN / 4); idx += 256) {
            const int pix = idx / (CIN / 4), c4 = (idx % (CIN / 4)) * 4;
            const int y = ty0 + pix / OP_TX, x = tx0 + pix % OP_TX;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < H && x < W) v = *reinterpret_cast<const float4*>(tokens + (((long long)b * H + y) * W + x) * ld + c4);
            *reinterpret_cast<float4*>(xs + pix * XS + c4) = v;   // full fp32: split into hi / lo at the fragment load
        }
        for (int idx = threadIdx.x; idx < 3 * OP_HY * OP_HX; idx += 256) {
            const int co = idx / (OP_HY * OP_HX), rem = idx % (OP_HY * OP_HX);
            const int hy = rem / OP_HX, hx = rem % OP_HX;
            const int y = ty0 + hy - 1, x = tx0 + hx - 1;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = dout[(((long long)b * 3 + co) * H + y) * W + x];
            dyf[idx] = v;
            if (hy >= 1 && hy <= OP_TY && hx >= 1 && hx <= OP_TX) {   // the tile's own pixels: bias gradient
                db0 += co == 0 ? v : 0.f;
                db1 += co == 1 ? v : 0.f;
                db2 += co == 2 ? v : 0.f;
            }
        }
        __syncthreads();

        // ---- dX: this warp's row of 16 pixels (rows g, g + 8 of the m16 tile) ----
        {
            const int ly = warp;
            uint32_t a[4][4], al[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const float* r0 = dyf + aoff[2 * ks] + ly * OP_HX + g;
                const float* r1 = dyf + aoff[2 * ks + 1] + ly * OP_HX + g;
                const float v[4] = {(avalid >> (2 * ks)) & 1u ? r0[0] : 0.f, (avalid >> (2 * ks)) & 1u ? r0[8] : 0.f,
                                    (avalid >> (2 * ks + 1)) & 1u ? r1[0] : 0.f, (avalid >> (2 * ks + 1)) & 1u ? r1[8] : 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a[ks][c] = f2tf32(v[c]);
                    al[ks][c] = __float_as_uint(v[c] - __uint_as_float(a[ks][c]));
                }
            }
            const int y = ty0 + ly;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int o0 = (ks * 8 + t) * XS + nt * 8 + g, o1 = o0 + 4 * XS;
                    const uint32_t bf[2] = {__float_as_uint(wm[o0]), __float_as_uint(wm[o1])};
                    const uint32_t bl[2] = {__float_as_uint(wl[o0]), __float_as_uint(wl[o1])};
                    mma_tf32_16x8x8(acc, al[ks], bf);
                    mma_tf32_16x8x8(acc, a[ks], bl);
                    mma_tf32_16x8x8(acc, a[ks], bf);
                }
                if (y < H) {
                    float* dp = dtokens + (((long long)b * H + y) * W + tx0) * CIN + nt * 8 + 2 * t;
                    if (tx0 + g < W) *reinterpret_cast<float2*>(dp + (long long)g * CIN) = make_float2(acc[0], acc[1]);
                    if (tx0 + g + 8 < W) *reinterpret_cast<float2*>(dp + (long long)(g + 8) * CIN) = make_float2(acc[2], acc[3]);
                }
            }
        }
        // ---- dW: rows = the 32 A columns, this warp's 8 input channels, contraction over its share of the pixels ----
        {
            constexpr int STEPS = NPIX / 8 / KSPLIT;
#pragma unroll 4
            for (int s8 = 0; s8 < STEPS; ++s8) {
                const int pix = (kh * STEPS + s8) * 8 + t;           // pixel of fragment column t (t + 4 is in the same row)
                const int poff = (pix / OP_TX) * OP_HX + pix % OP_TX;
                // 3xTF32 (hi / lo split of both operands): the weight gradients of the first and the last layer carry a
                // large share of the model's gradient norm, single-pass TF32 here was visible in the whole-model parity
                uint32_t bf[2], bl[2];
                split_hi_lo(xs[pix * XS + nt2 * 8 + g], bf[0], bl[0]);
                split_hi_lo(xs[(pix + 4) * XS + nt2 * 8 + g], bf[1], bl[1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    uint32_t a[4], al[4];
                    const bool v0 = (mvalid >> (2 * mt)) & 1u, v1 = (mvalid >> (2 * mt + 1)) & 1u;
                    split_hi_lo(v0 ? dyf[moff[2 * mt] + poff] : 0.f, a[0], al[0]);
                    split_hi_lo(v1 ? dyf[moff[2 * mt + 1] + poff] : 0.f, a[1], al[1]);
                    split_hi_lo(v0 ? dyf[moff[2 * mt] + poff + 4] : 0.f, a[2], al[2]);
                    split_hi_lo(v1 ? dyf[moff[2 * mt + 1] + poff + 4] : 0.f, a[3], al[3]);
                    mma_tf32_16x8x8(dwacc[mt], al, bf);
                    mma_tf32_16x8x8(dwacc[mt], a, bl);
                    mma_tf32_16x8x8(dwacc[mt], a, bf);
                }
            }
        }
    }
    // ---- per-CTA partial sums in the weight's own layout [co][ci][tap] (+ 3 bias sums), as the scalar kernel ----
    __syncthreads();
    float* red = smem;   // [32][CIN]
    for (int half = 0; half < KSPLIT; ++half) {
        if (kh == half) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int m = mt * 16 + g + (c >> 1) * 8, n = nt2 * 8 + 2 * t + (c & 1);
                    if (half == 0) red[m * CIN + n] = dwacc[mt][c];
                    else red[m * CIN + n] += dwacc[mt][c];
                }
        }
        __syncthreads();
    }
    float* part = partials + (long long)blockIdx.x * (27 * CIN + 3);
    for (int idx = threadIdx.x; idx < 27 * CIN; idx += 256) {
        const int co = idx / (9 * CIN), ci = (idx / 9) % CIN, tap = idx % 9;
        part[idx] = red[(co * 9 + tap) * CIN + ci];
    }
    db0 = warp_sum(db0);
    db1 = warp_sum(db1);
    db2 = warp_sum(db2);
    if (lane == 0) dbred[warp][0] = db0, dbred[warp][1] = db1, dbred[warp][2] = db2;
    __syncthreads();
    if (threadIdx.x < 3) {
        float sum = 0.f;
        for (int ww = 0; ww < 8; ++ww) sum += dbred[ww][threadIdx.x];
        part[27 * CIN + threadIdx.x] = sum;
    }
}

// InputProj (3 -> Cout, + LeakyReLU) on the tensor cores.  A[p][k] = img[ci][py + ky - 1][px + kx - 1], k = ci*9 + ky*3 + kx
// (27 columns padded to 32), read straight from the 3 x 10 x 18 halo tile of the image:
//   forward   T[p][co] = leaky(b[co] + sum_k A[p][k] Wm[k][co])      3xTF32 (hi/lo split of both operands): fp32-level
//   backward  dW[k][co] = sum_p A[p][k] dZ[p][co],  dZ = dT * leaky'(T)      3xTF32 as well (see the OutputProj kernel)
// The scalar kernels above run 864 FMAs per pixel on the CUDA cores (0.15 / 0.36 ms at B=16 256x256, 7x their HBM time).
template <int COUT>
__global__ void __launch_bounds__(256, 2) input_proj_fwd_mma_kernel(const float* __restrict__ img,
                                                                    const float* __restrict__ weight,
                                                                    const float* __restrict__ bias,
                                                                    float* __restrict__ tokens, int B, int H, int W,
                                                                    float slope, int tiles_x, int tiles_per_img) {
    uwr_pdl_enter();
    constexpr int NT = COUT / 8;
    __shared__ float ims[3 * OP_HY * OP_HX];
    // Wm[k][co], rows 27..31 zero, split once: hi = RN to TF32, lo = the exact remainder
    __shared__ uint32_t wh[32 * (COUT + OPM_XS)], wl[32 * (COUT + OPM_XS)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int idx = threadIdx.x; idx < 32 * COUT; idx += 256) {
        const int k = idx / COUT, co = idx % COUT;
        const float v = k < 27 ? weight[co * 27 + k] : 0.f;
        const uint32_t hi = f2tf32(v);
        wh[k * (COUT + OPM_XS) + co] = hi;
        wl[k * (COUT + OPM_XS) + co] = __float_as_uint(v - __uint_as_float(hi));
    }
    float2 bv[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bv[nt] = *reinterpret_cast<const float2*>(bias + nt * 8 + 2 * t);
    int aoff[8];
    unsigned avalid = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = (j >> 1) * 8 + t + (j & 1) * 4;
        const int kk = k < 27 ? k : 0;
        aoff[j] = (kk / 9) * (OP_HY * OP_HX) + ((kk % 9) / 3) * OP_HX + kk % 3;
        if (k < 27) avalid |= 1u << j;
    }
    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * OP_TY, tx0 = (tl % tiles_x) * OP_TX;
        __syncthreads();
        for (int idx = threadIdx.x; idx < 3 * OP_HY * OP_HX; idx += 256) {
            const int ci = idx / (OP_HY * OP_HX), rem = idx % (OP_HY * OP_HX);
            const int y = ty0 + rem / OP_HX - 1, x = tx0 + rem % OP_HX - 1;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = img[(((long long)b * 3 + ci) * H + y) * W + x];
            ims[idx] = v;
        }
        __syncthreads();
        const int ly = warp, y = ty0 + ly;
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 2 * ks + (c >> 1);
                const float v = (avalid >> j) & 1u ? ims[aoff[j] + ly * OP_HX + g + (c & 1) * 8] : 0.f;
                ah[ks][c] = f2tf32(v);
                al[ks][c] = __float_as_uint(v - __uint_as_float(ah[ks][c]));
            }
        if (y < H) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float acc[4] = {bv[nt].x, bv[nt].y, bv[nt].x, bv[nt].y};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {   // B fragments: k = ks*8 + t (+4), n = nt*8 + g
                    const int o0 = (ks * 8 + t) * (COUT + OPM_XS) + nt * 8 + g, o1 = o0 + 4 * (COUT + OPM_XS);
                    const uint32_t bh[2] = {wh[o0], wh[o1]}, bl[2] = {wl[o0], wl[o1]};
                    mma_tf32_16x8x8(acc, al[ks], bh);
                    mma_tf32_16x8x8(acc, ah[ks], bl);
                    mma_tf32_16x8x8(acc, ah[ks], bh);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = acc[c] > 0.f ? acc[c] : acc[c] * slope;
                float* tp = tokens + (((long long)b * H + y) * W + tx0) * COUT + nt * 8 + 2 * t;
                if (tx0 + g < W) *reinterpret_cast<float2*>(tp + (long long)g * COUT) = make_float2(acc[0], acc[1]);
                if (tx0 + g + 8 < W) *reinterpret_cast<float2*>(tp + (long long)(g + 8) * COUT) = make_float2(acc[2], acc[3]);
            }
        }
    }
}

template <int COUT>
__global__ void __launch_bounds__(256, 3) input_proj_bwd_mma_kernel(const float* __restrict__ dtokens,
                                                                    const float* __restrict__ tokens,
                                                                    const float* __restrict__ img,
                                                                    float* __restrict__ partials, int B, int H, int W,
                                                                    float slope, int tiles_x, int tiles_per_img) {
    uwr_pdl_enter();
    constexpr int ZS = COUT + OPM_XS;
    constexpr int NT = COUT / 8, KSPLIT = 8 / NT, NPIX = OP_TY * OP_TX;
    constexpr int C4 = COUT / 4;            // float4 groups per token; 256 % C4 == 0: a thread always stages the same group
    extern __shared__ __align__(16) float smem[];
    float* zs = smem;                       // [128][ZS]  dZ at the tile's pixels, full fp32
    float* ims = zs + NPIX * ZS;            // [3][10][18] image halo tile, full fp32
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    int moff[4];
    unsigned mvalid = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = (j >> 1) * 16 + g + (j & 1) * 8;
        const int mm = m < 27 ? m : 0;
        moff[j] = (mm / 9) * (OP_HY * OP_HX) + ((mm % 9) / 3) * OP_HX + mm % 3;
        if (m < 27) mvalid |= 1u << j;
    }
    const int nt2 = warp % NT, kh = warp / NT;
    float dwacc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int c = 0; c < 4; ++c) dwacc[mt][c] = 0.f;
    float4 dbs = make_float4(0.f, 0.f, 0.f, 0.f);   // bias gradient of this thread's channel group (full fp32 dZ)
    const int c4 = (threadIdx.x % C4) * 4;

    const int total = B * tiles_per_img;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tiles_per_img, tl = tile % tiles_per_img;
        const int ty0 = (tl / tiles_x) * OP_TY, tx0 = (tl % tiles_x) * OP_TX;
        __syncthreads();
        for (int idx = threadIdx.x; idx < NPIX * C4; idx += 256) {
            const int pix = idx / C4;
            const int y = ty0 + pix / OP_TX, x = tx0 + pix % OP_TX;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < H && x < W) {
                const long long tok = ((long long)b * H + y) * W + x;
                const float4 o = *reinterpret_cast<const float4*>(tokens + tok * COUT + c4);
                d = *reinterpret_cast<const float4*>(dtokens + tok * COUT + c4);
                d.x *= o.x > 0.f ? 1.f : slope; d.y *= o.y > 0.f ? 1.f : slope;
                d.z *= o.z > 0.f ? 1.f : slope; d.w *= o.w > 0.f ? 1.f : slope;
            }
            dbs.x += d.x; dbs.y += d.y; dbs.z += d.z; dbs.w += d.w;
            *reinterpret_cast<float4*>(zs + pix * ZS + c4) = d;   // full fp32: hi / lo split at the fragment load
        }
        for (int idx = threadIdx.x; idx < 3 * OP_HY * OP_HX; idx += 256) {
            const int ci = idx / (OP_HY * OP_HX), rem = idx % (OP_HY * OP_HX);
            const int y = ty0 + rem / OP_HX - 1, x = tx0 + rem % OP_HX - 1;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = img[(((long long)b * 3 + ci) * H + y) * W + x];
            ims[idx] = v;
        }
        __syncthreads();
        constexpr int STEPS = NPIX / 8 / KSPLIT;
#pragma unroll 4
        for (int s8 = 0; s8 < STEPS; ++s8) {
            const int pix = (kh * STEPS + s8) * 8 + t;
            const int poff = (pix / OP_TX) * OP_HX + pix % OP_TX;
            uint32_t bf[2], bl[2];   // 3xTF32, as the OutputProj weight gradient
            split_hi_lo(zs[pix * ZS + nt2 * 8 + g], bf[0], bl[0]);
            split_hi_lo(zs[(pix + 4) * ZS + nt2 * 8 + g], bf[1], bl[1]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t a[4], al[4];
                const bool v0 = (mvalid >> (2 * mt)) & 1u, v1 = (mvalid >> (2 * mt + 1)) & 1u;
                split_hi_lo(v0 ? ims[moff[2 * mt] + poff] : 0.f, a[0], al[0]);
                split_hi_lo(v1 ? ims[moff[2 * mt + 1] + poff] : 0.f, a[1], al[1]);
                split_hi_lo(v0 ? ims[moff[2 * mt] + poff + 4] : 0.f, a[2], al[2]);
                split_hi_lo(v1 ? ims[moff[2 * mt + 1] + poff + 4] : 0.f, a[3], al[3]);
                mma_tf32_16x8x8(dwacc[mt], al, bf);
                mma_tf32_16x8x8(dwacc[mt], a, bl);
                mma_tf32_16x8x8(dwacc[mt], a, bf);
            }
        }
    }
    // per-CTA partials [k][Cout], k = 27 -> bias (the layout input_proj_reduce_kernel sums)
    __syncthreads();
    float* red = smem;          // [32][COUT]
    float* dbr = smem + 32 * COUT;   // [256][4]
    for (int half = 0; half < KSPLIT; ++half) {
        if (kh == half) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int m = mt * 16 + g + (c >> 1) * 8, n = nt2 * 8 + 2 * t + (c & 1);
                    if (half == 0) red[m * COUT + n] = dwacc[mt][c];
                    else red[m * COUT + n] += dwacc[mt][c];
                }
        }
        __syncthreads();
    }
    *reinterpret_cast<float4*>(dbr + threadIdx.x * 4) = dbs;
    __syncthreads();
    float* part = partials + (long long)blockIdx.x * 28 * COUT;
    for (int idx = threadIdx.x; idx < 27 * COUT; idx += 256) part[idx] = red[idx];
    if (threadIdx.x < COUT) {
        float sum = 0.f;
        for (int j = 0; j < 256 / C4; ++j) sum += dbr[(j * C4 + threadIdx.x / 4) * 4 + threadIdx.x % 4];
        part[27 * COUT + threadIdx.x] = sum;
    }
}

__global__ void output_proj_reduce_kernel(const float* __restrict__ partials, float* __restrict__ dweight,
                                          float* __restrict__ dbias, int P, int Cin) {
    uwr_pdl_enter();
    const int n = 27 * Cin + 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += partials[(long long)p * n + idx];
    if (idx < 27 * Cin) dweight[idx] = s;
    else dbias[idx - 27 * Cin] = s;
}

// ------------------------------------------------------------------------------------------
__global__ void im2col_4x4s2_kernel(const float* __restrict__ x, long long ld, float* __restrict__ col, int B, int H,
                                    int W, int C, int rnd) {
    uwr_pdl_enter();
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const long long total = (long long)B * Ho * Wo * 16 * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int tap = (int)(r % 16);
        r /= 16;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const int y = 2 * oy - 1 + (tap >> 2), xx = 2 * ox - 1 + (tap & 3);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && xx >= 0 && xx < W)
            v = *reinterpret_cast<const float4*>(x + (((long long)b * H + y) * W + xx) * ld + c4 * 4);
        if (rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        reinterpret_cast<float4*>(col)[i] = v;
    }
}

__global__ void col2im_4x4s2_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int B, int H, int W,
                                    int C) {
    uwr_pdl_enter();
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const long long total = (long long)B * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int xx = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ky = ((y + 1) & 1) + 2 * a;
            const int oy = (y + 1 - ky) / 2;
            if (y + 1 - ky < 0 || oy >= Ho) continue;
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const int kx = ((xx + 1) & 1) + 2 * bb;
                const int ox = (xx + 1 - kx) / 2;
                if (xx + 1 - kx < 0 || ox >= Wo) continue;
                const float4 v = *reinterpret_cast<const float4*>(
                    dcol + ((((long long)b * Ho + oy) * Wo + ox) * 16 + ky * 4 + kx) * C + c4 * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(dx)[i] = acc;
    }
}

// dense 3x3 stride-1 pad-1 convolution as a GEMM: col[row][(ky*3+kx)*C + ci] = x[b][y+ky-1][x+kx-1][ci]
__global__ void im2col_3x3_kernel(const float* __restrict__ x, long long ld, float* __restrict__ col, int B, int H,
                                  int W, int C, int rnd) {
    uwr_pdl_enter();
    const int C4 = C / 4;
    const long long total = (long long)B * H * W * 9 * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int tap = (int)(r % 9);
        r /= 9;
        const int xx = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const int sy = y + tap / 3 - 1, sx = xx + tap % 3 - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sy >= 0 && sy < H && sx >= 0 && sx < W)
            v = *reinterpret_cast<const float4*>(x + (((long long)b * H + sy) * W + sx) * ld + c4 * 4);
        if (rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        reinterpret_cast<float4*>(col)[i] = v;
    }
}

// dx[b][y][x][ci] (+)= sum_taps dcol[(b, y-ky+1, x-kx+1)][tap*C + ci]
__global__ void col2im_3x3_kernel(const float* __restrict__ dcol, float* __restrict__ dx, long long ld_dx, int B, int H,
                                  int W, int C, int accumulate) {
    uwr_pdl_enter();
    const int C4 = C / 4;
    const long long total = (long long)B * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int xx = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int oy = y - (tap / 3) + 1, ox = xx - (tap % 3) + 1;
            if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
            const float4 v = *reinterpret_cast<const float4*>(
                dcol + ((((long long)b * H + oy) * W + ox) * 9 + tap) * C + c4 * 4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4* d = reinterpret_cast<float4*>(dx + (((long long)b * H + y) * W + xx) * ld_dx + c4 * 4);
        if (accumulate) {
            const float4 o = *d;
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        *d = acc;
    }
}

__global__ void pixel_scatter_2x2_kernel(const float* __restrict__ g, const float* __restrict__ bias,
                                         float* __restrict__ out, long long ld_out, int B, int H, int W, int Cout) {
    uwr_pdl_enter();
    const long long total = (long long)B * H * W * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long r = i / Cout;
        const int x = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        const float bv = bias ? bias[co] : 0.f;
        const long long o00 = (((long long)b * 2 * H + 2 * y) * 2 * W + 2 * x) * ld_out + co;
        out[o00] = v.x + bv;
        out[o00 + ld_out] = v.y + bv;
        out[o00 + 2 * W * ld_out] = v.z + bv;
        out[o00 + 2 * W * ld_out + ld_out] = v.w + bv;
    }
}

__global__ void pixel_gather_2x2_kernel(const float* __restrict__ dout, long long ld_dout, float* __restrict__ dg,
                                        int B, int H, int W, int Cout, int rnd) {
    uwr_pdl_enter();
    const long long total = (long long)B * H * W * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long r = i / Cout;
        const int x = (int)(r % W);
        r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const long long o00 = (((long long)b * 2 * H + 2 * y) * 2 * W + 2 * x) * ld_dout + co;
        float4 v;
        v.x = dout[o00];
        v.y = dout[o00 + ld_dout];
        v.z = dout[o00 + 2 * W * ld_dout];
        v.w = dout[o00 + 2 * W * ld_dout + ld_dout];
        if (rnd) v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));  // GEMM operand
        reinterpret_cast<float4*>(dg)[i] = v;
    }
}

__global__ void copy2d_kernel(const float* __restrict__ src, long long ld_src, float* __restrict__ dst,
                              long long ld_dst, long long rows, int cols4, int accumulate) {
    uwr_pdl_enter();
    const long long total = rows * cols4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols4;
        const int c = (int)(i % cols4) * 4;
        float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
        float4* d = reinterpret_cast<float4*>(dst + r * ld_dst + c);
        if (accumulate) {
            const float4 o = *d;
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *d = v;
    }
}

// column sums: grid (P, ceil(cols/32)), block (32, 8)
__global__ void colsum_partial_kernel(const float* __restrict__ x, long long ld, float* __restrict__ partials,
                                      long long rows, int cols) {
    uwr_pdl_enter();
    __shared__ float sh[8][33];
    const int c = blockIdx.y * 32 + threadIdx.x;
    float s = 0.f;
    if (c < cols) {
        const long long step = (long long)gridDim.x * 8;
        long long r = (long long)blockIdx.x * 8 + threadIdx.y;
        for (; r + 7 * step < rows; r += 8 * step) {  // 8 independent loads in flight per thread
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(x + (r + i * step) * ld + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
        }
        for (; r < rows; r += step) s += __ldg(x + r * ld + c);
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        partials[(long long)blockIdx.x * cols + c] = t;
    }
}
// out[c] = sum_p partials[p][c]: 32 columns x 32 row slices per CTA, fixed order (deterministic)
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ partials,
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

int ew_blocks(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)uwr_sm_count() * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// Precision knob (accuracy studies): UWR_BOUNDARY_TF32=0 keeps the scalar fp32 kernels for the boundary convolutions in tf32 mode
bool boundary_tensor_core() {
    static const bool on = [] { const char* e = getenv("UWR_BOUNDARY_TF32"); return !(e && e[0] == '0'); }();
    return on && uwr_round_outputs();
}

int persistent_ctas(int tiles, int per_sm = 2) {
    int p = per_sm * uwr_sm_count();
    if (p > tiles) p = tiles;
    return p < 1 ? 1 : p;
}

}  // namespace

extern "C" int uwr_input_proj_fwd(const float* img, const float* weight, const float* bias, float* tokens, int B,
                                  int H, int W, int Cin, int Cout, float slope, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(img && weight && bias && tokens, "uwr_input_proj_fwd: null pointer");
    UWR_REQUIRE(Cin >= 1 && Cin <= IP_MAXCIN && (Cout == 32 || Cout == 64), "uwr_input_proj_fwd: Cin<=4, Cout in {32,64}");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_input_proj_fwd: bad batch");
    if (Cin == 3 && boundary_tensor_core()) {   // tensor-core kernel (3xTF32: fp32-level); tf32x3 mode keeps the scalar one
        const int mx = uwr_cdiv(W, OP_TX), my = uwr_cdiv(H, OP_TY);
        const int P = persistent_ctas(B * mx * my, 2);
        if (Cout == 32) (void)uwr_launch_pdl(input_proj_fwd_mma_kernel<32>, dim3(P), dim3(256), 0, stream, img, weight, bias, tokens, B, H, W, slope, mx, mx * my);
        else (void)uwr_launch_pdl(input_proj_fwd_mma_kernel<64>, dim3(P), dim3(256), 0, stream, img, weight, bias, tokens, B, H, W, slope, mx, mx * my);
        UWR_CHECK_LAUNCH("input_proj_fwd_mma_kernel");
        return 0;
    }
    const int tx = uwr_cdiv(W, IP_TS), ty = uwr_cdiv(H, IP_TS);
    dim3 grid(tx * ty, 1, B);
    if (Cout == 32) (void)uwr_launch_pdl(input_proj_fwd_kernel<1>, dim3(grid), dim3(256), 0, stream, img, weight, bias, tokens, H, W, Cin, Cout, slope, tx);
    else (void)uwr_launch_pdl(input_proj_fwd_kernel<2>, dim3(grid), dim3(256), 0, stream, img, weight, bias, tokens, H, W, Cin, Cout, slope, tx);
    UWR_CHECK_LAUNCH("input_proj_fwd_kernel");
    return 0;
}

extern "C" size_t uwr_input_proj_bwd_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    const int tiles = B * uwr_cdiv(W, OP_TX) * uwr_cdiv(H, OP_TY);   // the tensor-core kernel's tiling: the larger count
    return (size_t)persistent_ctas(tiles, 3) * (Cin * 9 + 1) * Cout * sizeof(float);
}

extern "C" int uwr_input_proj_bwd(const float* dtokens, const float* tokens, const float* img, float* dweight,
                                  float* dbias, float* workspace, int B, int H, int W, int Cin, int Cout, float slope,
                                  uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dtokens && tokens && img && dweight && dbias && workspace, "uwr_input_proj_bwd: null pointer");
    UWR_REQUIRE(Cin >= 1 && Cin <= IP_MAXCIN && (Cout == 32 || Cout == 64), "uwr_input_proj_bwd: Cin<=4, Cout in {32,64}");
    if (Cin == 3 && boundary_tensor_core()) {
        const int mx = uwr_cdiv(W, OP_TX), my = uwr_cdiv(H, OP_TY);
        const int P = persistent_ctas(B * mx * my, 3);
        const int msmem = (OP_TY * OP_TX * (Cout + OPM_XS) + 3 * OP_HY * OP_HX) * (int)sizeof(float);
        if (Cout == 32) {
            (void)uwr_launch_pdl(input_proj_bwd_mma_kernel<32>, dim3(P), dim3(256), msmem, stream, dtokens, tokens, img, workspace, B, H, W, slope, mx, mx * my);
        } else {
            static bool configured = false;
            if (!configured) {
                UWR_CUDA(cudaFuncSetAttribute(input_proj_bwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
                configured = true;
            }
            (void)uwr_launch_pdl(input_proj_bwd_mma_kernel<64>, dim3(P), dim3(256), msmem, stream, dtokens, tokens, img, workspace, B, H, W, slope, mx, mx * my);
        }
        UWR_CHECK_LAUNCH("input_proj_bwd_mma_kernel");
        (void)uwr_launch_pdl(input_proj_reduce_kernel, dim3(uwr_cdiv(28 * Cout, 128)), dim3(128), 0, stream, workspace, dweight, dbias, P, Cin, Cout);
        UWR_CHECK_LAUNCH("input_proj_reduce_kernel");
        return 0;
    }
    const int tx = uwr_cdiv(W, IP_TS), ty = uwr_cdiv(H, IP_TS);
    const int P = persistent_ctas(B * tx * ty);
    if (Cout == 32)
        (void)uwr_launch_pdl(input_proj_bwd_kernel<1>, dim3(P), dim3(256), 0, stream, dtokens, tokens, img, workspace, B, H, W, Cin, Cout, slope, tx, tx * ty);
    else
        (void)uwr_launch_pdl(input_proj_bwd_kernel<2>, dim3(P), dim3(256), 0, stream, dtokens, tokens, img, workspace, B, H, W, Cin, Cout, slope, tx, tx * ty);
    UWR_CHECK_LAUNCH("input_proj_bwd_kernel");
    (void)uwr_launch_pdl(input_proj_reduce_kernel, dim3(uwr_cdiv((Cin * 9 + 1) * Cout, 128)), dim3(128), 0, stream, workspace, dweight, dbias, P, Cin, Cout);
    UWR_CHECK_LAUNCH("input_proj_reduce_kernel");
    return 0;
}

extern "C" int uwr_output_proj_fwd(const float* tokens, long long ld, const float* weight, const float* bias,
                                   const float* residual_img, float* out_img, int B, int H, int W, int Cin,
                                   uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(tokens && weight && bias && out_img, "uwr_output_proj_fwd: null pointer");
    UWR_REQUIRE((Cin == 32 || Cin == 64) && ld % 4 == 0, "uwr_output_proj_fwd: Cin in {32,64}, ld %% 4 == 0");
    UWR_REQUIRE(B > 0 && B <= 65535, "uwr_output_proj_fwd: bad batch");
    const int tx = uwr_cdiv(W, OP_TX), ty = uwr_cdiv(H, OP_TY);
    dim3 grid(tx * ty, 1, B);
    const int smem = OP_HY * OP_HX * Cin * (int)sizeof(float);
    if (Cin == 32) {
        (void)uwr_launch_pdl(output_proj_fwd_kernel<1>, dim3(grid), dim3(256), smem, stream, tokens, ld, weight, bias, residual_img, out_img, H, W, tx);
    } else {
        static bool configured = false;
        if (!configured) {
            UWR_CUDA(cudaFuncSetAttribute(output_proj_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = true;
        }
        (void)uwr_launch_pdl(output_proj_fwd_kernel<2>, dim3(grid), dim3(256), smem, stream, tokens, ld, weight, bias, residual_img, out_img, H, W, tx);
    }
    UWR_CHECK_LAUNCH("output_proj_fwd_kernel");
    return 0;
}

extern "C" size_t uwr_output_proj_bwd_workspace_bytes(int B, int H, int W, int Cin) {
    const int tiles = B * uwr_cdiv(W, OP_TX) * uwr_cdiv(H, OP_TY);
    return (size_t)persistent_ctas(tiles, 3) * (27 * Cin + 3) * sizeof(float);   // the tensor-core kernel runs 3 CTAs per SM
}

extern "C" int uwr_output_proj_bwd(const float* dout_img, const float* tokens, long long ld, const float* weight,
                                   float* dtokens, float* dweight, float* dbias, float* workspace, int B, int H, int W,
                                   int Cin, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dout_img && tokens && weight && dtokens && dweight && dbias && workspace, "uwr_output_proj_bwd: null pointer");
    UWR_REQUIRE((Cin == 32 || Cin == 64) && ld % 4 == 0, "uwr_output_proj_bwd: Cin in {32,64}, ld %% 4 == 0");
    const int tx = uwr_cdiv(W, OP_TX), ty = uwr_cdiv(H, OP_TY);
    const int P = persistent_ctas(B * tx * ty, boundary_tensor_core() ? 3 : 2);
    // tile + dY halo; the cross-warp reduction reuses the same buffer (8 * 9 * Cin floats)
    int smem = (OP_HY * OP_HX * Cin + 3 * OP_HY * OP_HX) * (int)sizeof(float);
    const int red = 8 * 9 * Cin * (int)sizeof(float);
    if (red > smem) smem = red;
    if (boundary_tensor_core()) {   // single-pass TF32 mode: both gradients on the tensor cores
        const int msmem = ((OP_TY * OP_TX + 64) * (Cin + OPM_XS) + 3 * OP_HY * OP_HX) * (int)sizeof(float);
        if (Cin == 32) {
            static bool configured = false;
            if (!configured) {
                UWR_CUDA(cudaFuncSetAttribute(output_proj_bwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
                configured = true;
            }
            (void)uwr_launch_pdl(output_proj_bwd_mma_kernel<32>, dim3(P), dim3(256), msmem, stream, dout_img, tokens, ld, weight, dtokens, workspace, B, H, W, tx, tx * ty);
        } else {
            static bool configured = false;
            if (!configured) {
                UWR_CUDA(cudaFuncSetAttribute(output_proj_bwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
                configured = true;
            }
            (void)uwr_launch_pdl(output_proj_bwd_mma_kernel<64>, dim3(P), dim3(256), msmem, stream, dout_img, tokens, ld, weight, dtokens, workspace, B, H, W, tx, tx * ty);
        }
        UWR_CHECK_LAUNCH("output_proj_bwd_mma_kernel");
    } else if (Cin == 32) {
        static bool configured = false;
        if (!configured) {
            UWR_CUDA(cudaFuncSetAttribute(output_proj_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = true;
        }
        (void)uwr_launch_pdl(output_proj_bwd_kernel<1>, dim3(P), dim3(256), smem, stream, dout_img, tokens, ld, weight, dtokens, workspace, B, H, W, tx, tx * ty);
    } else {
        static bool configured = false;
        if (!configured) {
            UWR_CUDA(cudaFuncSetAttribute(output_proj_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = true;
        }
        (void)uwr_launch_pdl(output_proj_bwd_kernel<2>, dim3(P), dim3(256), smem, stream, dout_img, tokens, ld, weight, dtokens, workspace, B, H, W, tx, tx * ty);
    }
    UWR_CHECK_LAUNCH("output_proj_bwd_kernel");
    (void)uwr_launch_pdl(output_proj_reduce_kernel, dim3(uwr_cdiv(27 * Cin + 3, 128)), dim3(128), 0, stream, workspace, dweight, dbias, P, Cin);
    UWR_CHECK_LAUNCH("output_proj_reduce_kernel");
    return 0;
}

extern "C" int uwr_im2col_4x4s2(const float* tokens, long long ld, float* col, int B, int H, int W, int C,
                                uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(tokens && col && C % 4 == 0 && ld % 4 == 0 && H % 2 == 0 && W % 2 == 0, "uwr_im2col_4x4s2: bad args");
    const long long total = (long long)B * (H / 2) * (W / 2) * 16 * (C / 4);
    (void)uwr_launch_pdl(im2col_4x4s2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, tokens, ld, col, B, H, W, C, uwr_round_outputs());
    UWR_CHECK_LAUNCH("im2col_4x4s2_kernel");
    return 0;
}

extern "C" int uwr_col2im_4x4s2(const float* dcol, float* dtokens, int B, int H, int W, int C, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dcol && dtokens && C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "uwr_col2im_4x4s2: bad args");
    const long long total = (long long)B * H * W * (C / 4);
    (void)uwr_launch_pdl(col2im_4x4s2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, dcol, dtokens, B, H, W, C);
    UWR_CHECK_LAUNCH("col2im_4x4s2_kernel");
    return 0;
}

extern "C" int uwr_im2col_3x3(const float* tokens, long long ld, float* col, int B, int H, int W, int C,
                              uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(tokens && col && C % 4 == 0 && ld % 4 == 0, "uwr_im2col_3x3: bad args");
    const long long total = (long long)B * H * W * 9 * (C / 4);
    (void)uwr_launch_pdl(im2col_3x3_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, tokens, ld, col, B, H, W, C, uwr_round_outputs());
    UWR_CHECK_LAUNCH("im2col_3x3_kernel");
    return 0;
}

extern "C" int uwr_col2im_3x3(const float* dcol, float* dtokens, long long ld, int B, int H, int W, int C,
                              int accumulate, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dcol && dtokens && C % 4 == 0 && ld % 4 == 0, "uwr_col2im_3x3: bad args");
    const long long total = (long long)B * H * W * (C / 4);
    (void)uwr_launch_pdl(col2im_3x3_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, dcol, dtokens, ld, B, H, W, C, accumulate);
    UWR_CHECK_LAUNCH("col2im_3x3_kernel");
    return 0;
}

extern "C" int uwr_pixel_scatter_2x2(const float* g, const float* bias, float* out, long long ld_out, int B, int H,
                                     int W, int Cout, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(g && out, "uwr_pixel_scatter_2x2: null pointer");
    const long long total = (long long)B * H * W * Cout;
    (void)uwr_launch_pdl(pixel_scatter_2x2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, g, bias, out, ld_out, B, H, W, Cout);
    UWR_CHECK_LAUNCH("pixel_scatter_2x2_kernel");
    return 0;
}

extern "C" int uwr_pixel_gather_2x2(const float* dout, long long ld_dout, float* dg, int B, int H, int W, int Cout,
                                    uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(dout && dg, "uwr_pixel_gather_2x2: null pointer");
    const long long total = (long long)B * H * W * Cout;
    (void)uwr_launch_pdl(pixel_gather_2x2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, dout, ld_dout, dg, B, H, W, Cout,
                                                                       uwr_round_outputs());
    UWR_CHECK_LAUNCH("pixel_gather_2x2_kernel");
    return 0;
}

// PixelShuffle(2) / PixelUnshuffle(2) on NHWC tokens (SpectralTransformer.py:151-158,191-198; block.py:107-153):
// channel co*4 + dy*2 + dx of pixel (y, x)  <->  channel co of pixel (2y+dy, 2x+dx); plain data movement, no rounding.
extern "C" int uwr_pixel_shuffle2(const float* in, float* out, long long ld_out, int B, int H, int W, int Cout,
                                  uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(in && out && ((uintptr_t)in & 15) == 0, "uwr_pixel_shuffle2: null / unaligned pointer");
    const long long total = (long long)B * H * W * Cout;
    (void)uwr_launch_pdl(pixel_scatter_2x2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, in, nullptr, out, ld_out, B, H, W, Cout);
    UWR_CHECK_LAUNCH("pixel_scatter_2x2_kernel");
    return 0;
}

extern "C" int uwr_pixel_unshuffle2(const float* in, long long ld_in, float* out, int B, int H, int W, int Cout,
                                    uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(in && out && ((uintptr_t)out & 15) == 0, "uwr_pixel_unshuffle2: null / unaligned pointer");
    const long long total = (long long)B * H * W * Cout;
    (void)uwr_launch_pdl(pixel_gather_2x2_kernel, dim3(ew_blocks(total, 256)), dim3(256), 0, stream, in, ld_in, out, B, H, W, Cout, 0);
    UWR_CHECK_LAUNCH("pixel_gather_2x2_kernel");
    return 0;
}

extern "C" int uwr_copy2d(const float* src, long long ld_src, float* dst, long long ld_dst, long long rows, int cols,
                          int accumulate, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(src && dst && cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, "uwr_copy2d: cols/ld must be multiples of 4");
    if (rows == 0 || cols == 0) return 0;
    (void)uwr_launch_pdl(copy2d_kernel, dim3(ew_blocks(rows * (cols / 4), 256)), dim3(256), 0, stream, src, ld_src, dst, ld_dst, rows, cols / 4, accumulate);
    UWR_CHECK_LAUNCH("copy2d_kernel");
    return 0;
}

extern "C" int uwr_colsum(const float* x, long long ld, float* out, float* workspace, long long rows, int cols,
                          uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UWR_REQUIRE(x && out && workspace && cols > 0 && rows > 0, "uwr_colsum: bad args");
    long long P = (rows + 63) / 64;
    const int cgroups = uwr_cdiv(cols, 32);
    long long cap = (4LL * uwr_sm_count() + cgroups - 1) / cgroups;
    if (cap > 1024) cap = 1024;
    if (P > cap) P = cap;
    if (P < 1) P = 1;
    (void)uwr_launch_pdl(colsum_partial_kernel, dim3(dim3((unsigned)P, cgroups)), dim3(dim3(32, 8)), 0, stream, x, ld, workspace, rows, cols);
    UWR_CHECK_LAUNCH("colsum_partial_kernel");
    (void)uwr_launch_pdl(colsum_final_kernel, dim3(uwr_cdiv(cols, 32)), dim3(1024), 0, stream, workspace, out, (int)P, cols);
    UWR_CHECK_LAUNCH("colsum_final_kernel");
    return 0;
}
