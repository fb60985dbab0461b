// Direct 3x3 (stride 1, pad 1) convolutions for the thin ends of SpectralTransformer and NewBigFRFN, where one side
// has 3 or 8 channels and a GEMM formulation would be all padding:
//
//   embed_conv_rgb 3 -> 16 (SpectralTransformer.py:217,255)      input_proj.proj.0  3 -> 8 (block.py:42-64)
//   output         8 -> 3  (SpectralTransformer.py:250,269)      output_proj.proj.2 8 -> 3 (block.py:67-88) [+ x]
//
// Either side is an NCHW image (B, C, H, W) or a token matrix (B*H*W, ld) (NHWC).  One thread = one output pixel,
// all output channels in registers, the 3x3xCin window read through L1 (neighbouring threads share 2/3 of it),
// weights in shared memory.  Algorithmic bytes: read Cin + write Cout floats per pixel -> HBM-bound.
// Backward: data gradient = the same kernel with flipped / transposed weights (prepared by the host);
// weight + bias gradient = per-CTA partial sums over 16x16 pixel tiles, then one deterministic reduce.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"

namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_TILE = 16;

template <int CIN, int COUT, bool IN_TOK, bool OUT_TOK>
__global__ void __launch_bounds__(CS_THREADS)
conv_small_fwd_kernel(const float* __restrict__ in, long long ld_in, const float* __restrict__ w,
                      const float* __restrict__ bias, const float* __restrict__ resid, float* __restrict__ out,
                      long long ld_out, int B, int H, int W, int rnd) {
    uwr_pdl_enter();
    __shared__ float ws[COUT * CIN * 9 + COUT];
    for (int i = threadIdx.x; i < COUT * CIN * 9; i += CS_THREADS) ws[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += CS_THREADS) ws[COUT * CIN * 9 + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const long long HW = (long long)H * W, total = (long long)B * HW;
    for (long long pix = (long long)blockIdx.x * CS_THREADS + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * CS_THREADS) {
        const long long b = pix / HW;
        const int rem = (int)(pix - b * HW), y = rem / W, x = rem - y * W;
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = ws[COUT * CIN * 9 + co];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= W) continue;
                float v[CIN];
                if (IN_TOK) {
                    const float* src = in + (b * HW + (long long)yy * W + xx) * ld_in;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) v[ci] = src[ci];
                } else {
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) v[ci] = in[((b * CIN + ci) * H + yy) * (long long)W + xx];
                }
#pragma unroll
                for (int co = 0; co < COUT; ++co)
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) acc[co] = fmaf(ws[(co * CIN + ci) * 9 + ky * 3 + kx], v[ci], acc[co]);
            }
        }
        if (OUT_TOK) {
            float* dst = out + pix * ld_out;
#pragma unroll
            for (int co = 0; co < COUT; ++co) dst[co] = rnd ? tf32_round(acc[co]) : acc[co];
        } else {
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const long long o = ((b * COUT + co) * H + y) * (long long)W + x;
                out[o] = acc[co] + (resid ? resid[o] : 0.f);
            }
        }
    }
}

// dW[co,ci,ky,kx] = sum_pix dout[pix,co] * in[pix + (ky-1, kx-1), ci];  db[co] = sum_pix dout[pix,co]
template <int CIN, int COUT, bool IN_TOK, bool DOUT_TOK>
__global__ void __launch_bounds__(CS_THREADS)
conv_small_wgrad_kernel(const float* __restrict__ in, long long ld_in, const float* __restrict__ dout, long long ld_dout,
                        float* __restrict__ partials, int B, int H, int W) {
    uwr_pdl_enter();
    constexpr int HT = CS_TILE + 2;
    constexpr int NOUT = COUT * CIN * 9 + COUT;
    constexpr int PER = (NOUT + CS_THREADS - 1) / CS_THREADS;
    __shared__ float ins[HT * HT * CIN];
    __shared__ float ds[CS_TILE * CS_TILE * COUT];
    const int tx_n = (W + CS_TILE - 1) / CS_TILE, ty_n = (H + CS_TILE - 1) / CS_TILE;
    const long long tiles = (long long)B * tx_n * ty_n;
    const long long HW = (long long)H * W;
    float acc[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) acc[k] = 0.f;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long b = tile / (tx_n * ty_n);
        const int tr = (int)(tile - b * tx_n * ty_n), y0 = (tr / tx_n) * CS_TILE, x0 = (tr % tx_n) * CS_TILE;
        __syncthreads();
        for (int i = threadIdx.x; i < HT * HT * CIN; i += CS_THREADS) {
            const int ci = i % CIN, p = i / CIN, yy = y0 + p / HT - 1, xx = x0 + p % HT - 1;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                v = IN_TOK ? in[(b * HW + (long long)yy * W + xx) * ld_in + ci] : in[((b * CIN + ci) * H + yy) * (long long)W + xx];
            ins[i] = v;
        }
        for (int i = threadIdx.x; i < CS_TILE * CS_TILE * COUT; i += CS_THREADS) {
            const int co = i % COUT, p = i / COUT, yy = y0 + p / CS_TILE, xx = x0 + p % CS_TILE;
            float v = 0.f;
            if (yy < H && xx < W)
                v = DOUT_TOK ? dout[(b * HW + (long long)yy * W + xx) * ld_dout + co]
                             : dout[((b * COUT + co) * H + yy) * (long long)W + xx];
            ds[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int o = threadIdx.x + k * CS_THREADS;
            if (o >= NOUT) break;
            float s = 0.f;
            if (o < COUT * CIN * 9) {
                const int co = o / (CIN * 9), r = o - co * CIN * 9, ci = r / 9, tap = r - ci * 9, ky = tap / 3, kx = tap - ky * 3;
                for (int py = 0; py < CS_TILE; ++py)
#pragma unroll 4
                    for (int px = 0; px < CS_TILE; ++px)
                        s = fmaf(ds[(py * CS_TILE + px) * COUT + co], ins[((py + ky) * HT + px + kx) * CIN + ci], s);
            } else {
                const int co = o - COUT * CIN * 9;
                for (int p = 0; p < CS_TILE * CS_TILE; ++p) s += ds[p * COUT + co];
            }
            acc[k] += s;
        }
    }
    float* part = partials + (long long)blockIdx.x * NOUT;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int o = threadIdx.x + k * CS_THREADS;
        if (o < NOUT) part[o] = acc[k];
    }
}

__global__ void conv_small_reduce_kernel(const float* __restrict__ partials, int nparts, int nw, int nb,
                                         float* __restrict__ dweight, float* __restrict__ dbias) {
    uwr_pdl_enter();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nw + nb) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partials[(long long)p * (nw + nb) + o];
    if (o < nw) dweight[o] = s;
    else if (dbias) dbias[o - nw] = s;
}

int wgrad_ctas(int B, int H, int W) {
    const long long tiles = (long long)B * uwr_cdiv(H, CS_TILE) * uwr_cdiv(W, CS_TILE);
    const long long cap = 4LL * uwr_sm_count();
    return (int)(tiles < cap ? tiles : cap);
}

template <int CIN, int COUT, bool IN_TOK, bool OUT_TOK>
int launch_fwd(const uwr_conv_small_desc* d, cudaStream_t stream) {
    const long long total = (long long)d->B * d->H * d->W;
    long long grid = (total + CS_THREADS - 1) / CS_THREADS;
    const long long cap = 16LL * uwr_sm_count();
    if (grid > cap) grid = cap;
    (void)uwr_launch_pdl(conv_small_fwd_kernel<CIN, COUT, IN_TOK, OUT_TOK>, dim3((int)grid), dim3(CS_THREADS), 0, stream, 
        d->in, d->ld_in, d->weight, d->bias, d->residual_img, d->out, d->ld_out, d->B, d->H, d->W,
        (OUT_TOK && d->round_out) ? uwr_round_outputs() : 0);
    UWR_CHECK_LAUNCH("conv_small_fwd_kernel");
    return 0;
}

template <int CIN, int COUT, bool IN_TOK, bool DOUT_TOK>
int launch_wgrad(const uwr_conv_small_desc* d, const float* dout, long long ld_dout, float* dweight, float* dbias,
                 float* ws, cudaStream_t stream) {
    const int ctas = wgrad_ctas(d->B, d->H, d->W);
    (void)uwr_launch_pdl(conv_small_wgrad_kernel<CIN, COUT, IN_TOK, DOUT_TOK>, dim3(ctas), dim3(CS_THREADS), 0, stream, d->in, d->ld_in, dout, ld_dout, ws,
                                                                                         d->B, d->H, d->W);
    UWR_CHECK_LAUNCH("conv_small_wgrad_kernel");
    const int nw = COUT * CIN * 9;
    (void)uwr_launch_pdl(conv_small_reduce_kernel, dim3(uwr_cdiv(nw + COUT, 128)), dim3(128), 0, stream, ws, ctas, nw, COUT, dweight, dbias);
    UWR_CHECK_LAUNCH("conv_small_reduce_kernel");
    return 0;
}

int check_desc(const uwr_conv_small_desc* d, const char* who) {
    UWR_REQUIRE(d && d->in && d->weight, "%s: null pointer", who);
    UWR_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0, "%s: bad extent", who);
    const bool ok = (d->in_tokens == 0 && d->out_tokens == 1 && d->Cin == 3 && (d->Cout == 8 || d->Cout == 16)) ||
                    (d->in_tokens == 1 && d->out_tokens == 0 && d->Cin == 8 && d->Cout == 3);
    UWR_REQUIRE(ok, "%s: served shapes are image 3 -> tokens 8|16 and tokens 8 -> image 3 (got %d -> %d, layouts %d -> %d)", who,
                d->Cin, d->Cout, d->in_tokens, d->out_tokens);
    return 0;
}

}  // namespace

extern "C" int uwr_conv3x3_small_fwd(const uwr_conv_small_desc* d, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_desc(d, "uwr_conv3x3_small_fwd")) return rc;
    UWR_REQUIRE(d->out, "uwr_conv3x3_small_fwd: null output");
    if (d->Cin == 3 && d->Cout == 8) return launch_fwd<3, 8, false, true>(d, stream);
    if (d->Cin == 3 && d->Cout == 16) return launch_fwd<3, 16, false, true>(d, stream);
    return launch_fwd<8, 3, true, false>(d, stream);
}

extern "C" size_t uwr_conv3x3_small_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout) {
    return (size_t)wgrad_ctas(B, H, W) * ((size_t)Cout * Cin * 9 + Cout) * sizeof(float);
}

// dout has the layout of the forward OUTPUT (out_tokens / ld_dout); d->out and d->residual_img are ignored.
extern "C" int uwr_conv3x3_small_wgrad(const uwr_conv_small_desc* d, const float* dout, long long ld_dout, float* dweight,
                                       float* dbias, float* workspace, uwr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_desc(d, "uwr_conv3x3_small_wgrad")) return rc;
    UWR_REQUIRE(dout && dweight && workspace, "uwr_conv3x3_small_wgrad: null pointer");
    if (d->Cin == 3 && d->Cout == 8) return launch_wgrad<3, 8, false, true>(d, dout, ld_dout, dweight, dbias, workspace, stream);
    if (d->Cin == 3 && d->Cout == 16) return launch_wgrad<3, 16, false, true>(d, dout, ld_dout, dweight, dbias, workspace, stream);
    return launch_wgrad<8, 3, true, false>(d, dout, ld_dout, dweight, dbias, workspace, stream);
}
