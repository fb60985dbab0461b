/*
 * uwr_b200 — C ABI of the B200-native hot path for the restoration transformers of
 * KarthikSundar2002/Underwater-Image-Restoration (AST / SpectralTransformer / NewBigFRFN
 * forward + backward, their losses and the optimizer step).
 *
 * Conventions (SURVEY.md §8b "New C-ABI"):
 *   - every pointer is a DEVICE pointer valid on `stream` unless it says "host";
 *   - tensors are dense fp32 unless stated; "tokens" means (B, L=H*W, C) row-major, i.e. NHWC;
 *   - no allocation, no synchronisation, no global state inside a call; scratch comes from the
 *     caller (`*_workspace_bytes` tells how much);
 *   - return 0 on success, negative on error; uwr_last_error() gives the message.
 *
 * The reference is pure PyTorch-eager (no FFI of its own); each entry cites the reference
 * Python that it replaces (paths relative to the reference root).
 */
#ifndef UWR_B200_H
#define UWR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* uwr_stream_t; /* == cudaStream_t */

/* ---- library ---------------------------------------------------------------------------- */
const char* uwr_last_error(void);
int uwr_abi_version(void);
int uwr_device_sm_count(void);
unsigned long long uwr_launch_count(void); /* kernels launched by this library so far */

/* ---- GEMM (nn.Linear fwd/bwd: AST.py:47-48,104,297,302,332,337; block.py:158-160) --------
 * C[M,N] = epilogue( opA(A)[M,K] * opB(B)[K,N] ), TF32 tensor-core math (operands rounded to
 * nearest tf32 in-kernel), fp32 accumulate.
 *   a_km = 0: A stored [M][K] (lda = row stride)       a_km = 1: A stored [K][M]
 *   b_nk = 1: B stored [N][K] (nn.Linear weight)       b_nk = 0: B stored [K][N]
 *   B2/bias2/n_split: optional second weight segment: rows >= n_split of the STORED B come from
 *                     B2 (to_q | to_kv fused projection, AST.py:59-60): output columns when
 *                     b_nk = 1, contraction rows when b_nk = 0.
 *   epilogue: v = acc + bias[n]; v *= rowscale[m / rows_per_group] (if rowscale);
 *             UWR_EPI_RESID: v += R[m,n];  UWR_EPI_MUL_DGELU: v *= gelu'(R[m,n]);  UWR_EPI_MUL: v *= R[m,n].
 *   a_km = 1 (weight gradient, contraction over tokens): rowscale indexes the K rows
 *             (DropPath scale of the incoming gradient), colsum[M] = sum_k s_k A[k][m]
 *             (bias gradient) is produced when non-NULL, and the contraction is split over
 *             CTAs through `workspace` (uwr_gemm_workspace_bytes).
 */
enum { UWR_EPI_NONE = 0, UWR_EPI_RESID = 1, UWR_EPI_MUL_DGELU = 2, UWR_EPI_MUL = 3 };

typedef struct {
    const float* A;
    long long lda;
    int a_km;
    const float* B;
    long long ldb;
    int b_nk;
    const float* B2;
    int n_split;
    float* C;
    long long ldc;
    int M, N, K;
    const float* bias;
    const float* bias2;
    int epilogue;
    const float* R;
    long long ldr;
    const float* rowscale;
    int rows_per_group;
    float* colsum;
    float* workspace;
    size_t workspace_bytes;
    int round_out; /* store C rounded to TF32 (it feeds another tensor-core GEMM) */
    /* fp16 storage of the 4C-wide LeFF tensors (tcgen05 path only; DESIGN.md §3 "half storage"): a 10-bit mantissa
     * is exactly what a TF32 operand keeps, so for O(1)-range activations fp16 halves the HBM bytes at TF32 accuracy.
     * c_half: C is __half[M][ldc] (ldc in halves, values clamped to +-65504);
     * r_half: R (UWR_EPI_MUL multiplier / UWR_EPI_RESID residual) is __half[M][ldr]. */
    int c_half;
    int r_half;
} uwr_gemm_desc;

size_t uwr_gemm_workspace_bytes(int M, int N, int K, int a_km);
/* passes = 1: single TF32 product (default, ~4e-4 relative per GEMM);
 * passes = 3: error-compensated 3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi), fp32-level accuracy. */
int uwr_set_gemm_precision(int passes);
int uwr_get_gemm_precision(void);
int uwr_gemm_tf32(const uwr_gemm_desc* d, uwr_stream_t stream);

/* Blackwell-native path for the same descriptor: TMA (128B swizzle) -> smem ring -> tcgen05.mma
 * kind::tf32 with TMEM accumulators -> tcgen05.ld epilogue, persistent warp-specialised CTAs.
 * Operands must already be TF32-rounded (the tensor core truncates): producers round at store,
 * weights via uwr_round_tf32_tensors.  `_supported` tells whether a descriptor is served here
 * (no segmented weights / colsum / k-scale; MN-major widths in multiples of 32). */
int uwr_gemm_tcgen05_supported(const uwr_gemm_desc* d);
/* Thread-block cluster pairs with TMA multicast of the B tile for the tensor-bound shapes (arithmetic intensity >= 96
 * flop/B, N tile >= 128): 0 = off, 1 = auto (default: every eligible layout; +1 % on the training step),
 * 2 = the weight-gradient (TN) layout only. */
int uwr_set_gemm_cluster(int mode);
/* Programmatic dependent launch of the library's kernels (the next grid is launched while the running one drains and
 * blocks in griddepcontrol.wait until it has completed; results are those of plain stream order): 0 = plain launches (default),
 * 1 = on (also UWR_PDL=1 in the environment).  Pays for launch-bound steps (SpectralTransformer +2 %), not for AST. */
int uwr_set_pdl(int on);
size_t uwr_gemm_tcgen05_workspace_bytes(int M, int N, int K, int a_km);
int uwr_gemm_tcgen05(const uwr_gemm_desc* d, uwr_stream_t stream);

/* ---- convolutions on token (NHWC) tensors as implicit GEMMs on the tcgen05 kernel ------------------------------
 * Replaces im2col + GEMM for the dense convolutions between scales: Downsample Conv4x4 s2 p1 (AST.py:408-424), the 3x3
 * convolutions of NewBigFRFN (block.py:42-153) and SpectralTransformer (SpectralTransformer.py:133-159).  The im2col
 * matrix is never written: x (B*H*W, ld_x) is addressed by a 4-D (C, W, H, B) tensor map and every K chunk (32 channels of
 * one tap) is one TMA box at the tap's offset; out-of-image pixels are zero-filled by TMA (= the padding), stride 2 uses
 * the map's element strides.  Operands must be TF32-rounded (as for uwr_gemm_tcgen05).
 *   mode 0: y[B*OH*OW, Cout] = im2col(x) . w^T (+ bias)      w: (Cout, kh*kw*Cin), K index = (ky, kx, ci)
 *           (the data gradient of a stride-1 convolution is the same call on dy with the flipped, transposed weights)
 *   mode 1: dw[Cout, kh*kw*Cin] = dy^T . im2col(x)            dy: (B*OH*OW, ld_dy); contraction split across CTAs
 *   mode 2: dw^T[kh*kw*Cin, Cout] = im2col(x)^T . dy          the same gradient, transposed: the im2col view is the
 *           128-row operand, so a thin Cout costs a thin N tile instead of padding the M tile (5x fewer MMA cycles
 *           at Cout = 32, 3x3)
 * Served when Cin % 32 == 0 and OH*OW is a multiple of 128 (mode 0) / 32 (mode 1) with power-of-two or box-multiple
 * widths (see `_supported`); callers fall back to uwr_im2col_* + uwr_gemm_tcgen05 otherwise. */
typedef struct {
    int mode;
    const float* x;
    long long ld_x;
    int B, H, W, Cin;
    int kh, kw, stride, pad;
    int Cout;
    const float* w;      /* mode 0 */
    const float* bias;   /* mode 0, optional */
    float* y;            /* mode 0 */
    long long ld_y;
    int round_out;       /* mode 0: round y to TF32 at the store (it feeds another tensor-core product) */
    const float* dy;     /* mode 1 */
    long long ld_dy;
    float* dw;           /* mode 1: dense (Cout, kh*kw*Cin); mode 2: dense (kh*kw*Cin, Cout) */
    float* workspace;    /* modes 1, 2: uwr_convgemm_tcgen05_workspace_bytes */
    size_t workspace_bytes;
} uwr_convgemm_desc;
int uwr_convgemm_tcgen05_supported(const uwr_convgemm_desc* d);
size_t uwr_convgemm_tcgen05_workspace_bytes(const uwr_convgemm_desc* d);
int uwr_convgemm_tcgen05(const uwr_convgemm_desc* d, uwr_stream_t stream);

/* TF32 operand preparation (tcgen05 truncates; operands are rounded to nearest where produced).
 * In single-pass mode (uwr_set_gemm_precision(1)) the producers of GEMM operands — layernorm_fwd,
 * window_attn_fwd/bwd, dwconv_gelu_fwd (h2) / _bwd (du), im2col — round at their stores.
 * uwr_round_tf32_tensors: multi-tensor copy (+rounding when do_round) over device pointer tables
 * (offsets has n_tensors+1 entries); uwr_scale_round: dst[rows][cols] = tf32(rowscale[r/rpg]*src). */
int uwr_round_tf32_tensors(const float* const* src, float* const* dst, const long long* offsets,
                           int n_tensors, long long total_elems, int do_round, uwr_stream_t stream);
int uwr_scale_round(const float* src, long long ld_src, float* dst, long long rows, int cols,
                    const float* rowscale, int rows_per_group, int do_round, uwr_stream_t stream);
/* same, plus colsum[cols] = column sums of dst (bias gradient of the consuming Linear);
 * cols/4 must divide 256; workspace >= 1024*cols floats */
int uwr_scale_round_colsum(const float* src, long long ld_src, float* dst, long long rows, int cols,
                           const float* rowscale, int rows_per_group, int do_round, float* colsum,
                           float* workspace, uwr_stream_t stream);

/* ---- LayerNorm over C (nn.LayerNorm eps 1e-5: AST.py:521,534,593,622) ------------------- */
int uwr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y,
                      float* mean, float* rstd, long long rows, int C, float eps,
                      uwr_stream_t stream);
/* dx = dres (optional pass-through residual gradient) + LN backward; dgamma/dbeta reduced
 * deterministically through `partials` (uwr_layernorm_bwd_workspace_bytes). */
size_t uwr_layernorm_bwd_workspace_bytes(long long rows, int C);
int uwr_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                      const float* rstd, const float* dres, float* dx, float* dgamma,
                      float* dbeta, float* partials, long long rows, int C, uwr_stream_t stream);
/* uwr_layernorm_bwd that ALSO emits the GEMM operand of the next backward function: ds_out = tf32(rowscale[row /
 * rows_per_group] * dx) (DropPath-scaled; rounded in single-pass TF32 mode only) and ds_colsum = column sums of
 * ds_out (bias gradient of the Linear consuming it) -- what uwr_scale_round_colsum would compute in a separate pass.
 * Served for C in {16,32,64,128,256,512} (uwr_layernorm_bwd_ds_supported); 16-byte aligned pointers. */
int uwr_layernorm_bwd_ds_supported(long long rows, int C);
size_t uwr_layernorm_bwd_ds_workspace_bytes(long long rows, int C);
int uwr_layernorm_bwd_ds(const float* dy, const float* x, const float* gamma, const float* mean,
                         const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                         const float* ds_rowscale, int ds_rows_per_group, float* ds_out, float* ds_colsum,
                         float* workspace, long long rows, int C, uwr_stream_t stream);

/* ---- adaptive sparse window attention (WindowAttention_sparse.forward, AST.py:187-219;
 *      roll / window_partition / window_reverse, AST.py:377-402,596-618; shift mask 568-588) --
 * qkv: tokens (B, H*W, ld_qkv) with q at column q_off + h*HD, k at k_off + h*HD, v at
 * v_off + h*HD.  kv_ptr may differ from q_ptr (cross attention, block.py:185-188).
 * out: tokens (B, H*W, ld_out) written at column h*HD in ORIGINAL token order (the cyclic
 * shift and the 8x8 window partition are address arithmetic).  P = w0*softmax(S) + w1*relu(S)^2
 * with (w0,w1) = softmax(w_param) evaluated on the device.  HD in {8,16,32,64,128}.
 */
typedef struct {
    const float* q;
    long long ld_q;
    int q_off;
    const float* kv;
    long long ld_kv;
    int k_off;
    int v_off;
    const float* bias_table; /* (225, heads) */
    const float* w_param;    /* (2,) raw parameter; NULL => plain softmax attention */
    int B, H, W, heads, head_dim, shift;
    float scale;
    /* q, k, v (and dout in the backward) are exactly TF32-representable (their producing GEMM epilogues rounded them,
     * uwr_gemm_desc.round_out): the products are then exact in one tensor-core pass and the kernels skip the 3xTF32
     * hi/lo compensation they otherwise apply to full-fp32 operands. */
    int operands_rounded;
    /* backward only, both or neither: column sums of dq and of dk | dv (the bias gradients of the q / kv projections,
     * AST.py:139-157), written at the same column offsets as dq_buf / dkv_buf use (q_off + h*HD, k_off + h*HD,
     * v_off + h*HD; for a packed q|k|v buffer pass the same 3C vector twice).  NULL = not wanted.  Accumulated inside
     * the backward kernel, in a fixed order (deterministic), instead of one more pass over dq | dk | dv. */
    float* dq_colsum;
    float* dkv_colsum;
} uwr_attn_desc;

int uwr_window_attn_fwd(const uwr_attn_desc* d, float* out, long long ld_out, uwr_stream_t stream);
/* Forward on the tcgen05 tensor cores (head_dim 32, shift in {0, 4}, an even number of windows per image: two
 * windows per 128-row MMA tile, Q/K/V by TMA, scores / outputs in TMEM).  mode 0 = never, 1 = whenever eligible,
 * 2 = auto (default): when d->operands_rounded, where it is the faster kernel (0.279 vs 0.285 ms at 32 768 tiles;
 * with full-fp32 operands mma.sync wins, 0.316 vs 0.347 ms).  Both are parity-tested. */
int uwr_set_attn_tcgen05(int mode);
size_t uwr_window_attn_bwd_workspace_bytes(const uwr_attn_desc* d);
/* dq/dk/dv are written with the same (ld, offset) addressing as q/k/v into dq_buf/dkv_buf.
 * dbias_table (225,heads) and dw (2,) are overwritten. */
int uwr_window_attn_bwd(const uwr_attn_desc* d, const float* dout, long long ld_dout,
                        float* dq_buf, float* dkv_buf, float* dbias_table, float* dw,
                        float* workspace, uwr_stream_t stream);

/* ---- depthwise 3x3 conv on tokens with fused GELUs (LeFF / FRFN: AST.py:299-301,318,
 *      334-336,365-367) ---------------------------------------------------------------------
 * u: tokens (B,H,W,ld_u) pre-activation of linear1, channels [0,Ch) are convolved:
 *   v  = dwconv3x3(gelu(u)) + bias        (saved for backward when v != NULL; with v_is_dgelu the
 *                                          buffer receives gelu'(v) instead, which is all the LeFF
 *                                          backward needs: the GEMM epilogue UWR_EPI_MUL consumes it)
 *   h2 = gelu(v)                          (mode 0, LeFF)
 *   h2 = gelu(v) * gelu(u[:, Ch + c])     (mode 1, FRFN gate; u has 2*Ch channels)
 *   h2 = dwconv3x3(u) [+ bias]            (mode 2, plain depthwise conv: MDTA/GDFN,
 *                                          SpectralTransformer.py:82,89,123; bias may be NULL)
 */
int uwr_dwconv_gelu_fwd(const float* u, long long ld_u, const float* weight /*(Ch,1,3,3)*/,
                        const float* bias, float* v, float* h2, int B, int H, int W, int Ch,
                        int mode, int v_is_dgelu, uwr_stream_t stream);
/* dv = dh2 * gelu'(v) [* gelu(u2)] for callers that do not fuse it into the producing GEMM
 * (UWR_EPI_MUL_DGELU); mode 1 (FRFN) also writes du[:, Ch:2Ch] = dh2 * gelu(v) * gelu'(u2). */
int uwr_gelu_gate_bwd(const float* dh2, const float* u, long long ld_u, const float* v, float* dv,
                      float* du, long long rows, int Ch, int mode, uwr_stream_t stream);
/* GDFN gate (SpectralTransformer.py:126-129): out = gelu(t[:, :h]) * t[:, h:2h]; t has row stride ld >= 2h,
 * dt the same layout; h % 4 == 0. */
/* fp16 storage of the two 4C-wide LeFF tensors that are not tensor-core operands: u (linear1 output,
 * written by uwr_gemm_tcgen05 with c_half = 1) and gelu'(v) (read back by the linear2 data-gradient epilogue with
 * r_half = 1).  H, W multiples of 16, Ch of 32, single-pass (tf32) mode.  h2 and du stay fp32 (GEMM operands). */
int uwr_dwconv_half_supported(int H, int W, int Ch);
int uwr_dwconv_gelu_fwd_half(const void* u_half, const float* weight, const float* bias, void* dgelu_half,
                             float* h2, int B, int H, int W, int Ch, uwr_stream_t stream);
int uwr_dwconv_gelu_bwd_half(const float* dv, const void* u_half, const float* weight, float* du, float* dweight,
                             float* dbias, float* du_colsum, float* workspace, int B, int H, int W, int Ch,
                             uwr_stream_t stream);
int uwr_gelu_mul_fwd(const float* t, long long ld, float* out /*(rows,h)*/, long long rows, int h,
                     uwr_stream_t stream);
int uwr_gelu_mul_bwd(const float* dout /*(rows,h)*/, const float* t, long long ld, float* dt,
                     long long rows, int h, uwr_stream_t stream);
size_t uwr_dwconv_gelu_bwd_workspace_bytes(int B, int H, int W, int Ch);
/* dv (gradient w.r.t. the conv output v) -> du[:, :Ch] (row stride ld_u), dweight (Ch,1,3,3),
 * dbias (Ch); only dv needs a halo.  du_colsum (Ch, optional) = column sums of du, i.e. the bias
 * gradient of the Linear that produced u (saves a pass over du). */
int uwr_dwconv_gelu_bwd(const float* dv, const float* u, long long ld_u, const float* weight,
                        float* du, float* dweight, float* dbias, float* du_colsum, float* workspace, int B, int H,
                        int W, int Ch, int plain /* 1: no GELU around the conv (mode 2) */,
                        uwr_stream_t stream);

/* ---- convolutions at the model boundary and between scales ---------------------------------
 * InputProj  (AST.py:447-466): Conv3x3(3->Cout)+LeakyReLU(slope) NCHW image -> tokens.
 * OutputProj (AST.py:470-493, 920-921): tokens -> Conv3x3(Cin->3) [+ image residual] -> NCHW.
 * Downsample (AST.py:408-424): Conv4x4 s2 p1 as im2col (K order ky,kx,ci) + uwr_gemm_tf32.
 * Upsample   (AST.py:428-443): ConvTranspose2x2 s2 as uwr_gemm_tf32 + pixel scatter.
 */
int uwr_input_proj_fwd(const float* img, const float* weight, const float* bias, float* tokens,
                       int B, int H, int W, int Cin, int Cout, float slope, uwr_stream_t stream);
size_t uwr_input_proj_bwd_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int uwr_input_proj_bwd(const float* dtokens, const float* tokens, const float* img,
                       float* dweight, float* dbias, float* workspace, int B, int H, int W,
                       int Cin, int Cout, float slope, uwr_stream_t stream);
int uwr_output_proj_fwd(const float* tokens, long long ld, const float* weight, const float* bias,
                        const float* residual_img, float* out_img, int B, int H, int W, int Cin,
                        uwr_stream_t stream);
size_t uwr_output_proj_bwd_workspace_bytes(int B, int H, int W, int Cin);
int uwr_output_proj_bwd(const float* dout_img, const float* tokens, long long ld,
                        const float* weight, float* dtokens, float* dweight, float* dbias,
                        float* workspace, int B, int H, int W, int Cin, uwr_stream_t stream);
int uwr_im2col_4x4s2(const float* tokens, long long ld, float* col, int B, int H, int W, int C,
                     uwr_stream_t stream);
int uwr_col2im_4x4s2(const float* dcol, float* dtokens, int B, int H, int W, int C,
                     uwr_stream_t stream);
/* g: (B*H*W, Cout*4) GEMM result with column (co,dy,dx); out tokens (B,2H,2W,ld_out). */
/* dense 3x3 s1 p1 conv as GEMM: col (B*H*W, 9*C) with K order (ky,kx,ci); col2im gathers back */
int uwr_im2col_3x3(const float* tokens, long long ld, float* col, int B, int H, int W, int C,
                   uwr_stream_t stream);
int uwr_col2im_3x3(const float* dcol, float* dtokens, long long ld, int B, int H, int W, int C,
                   int accumulate, uwr_stream_t stream);
int uwr_pixel_scatter_2x2(const float* g, const float* bias, float* out, long long ld_out, int B,
                          int H, int W, int Cout, uwr_stream_t stream);
int uwr_pixel_gather_2x2(const float* dout, long long ld_dout, float* dg, int B, int H, int W,
                         int Cout, uwr_stream_t stream);
/* strided 2-D copy / accumulate: dst[r, 0:cols] (op)= src[r, 0:cols]  (skip concat, AST.py:904) */
int uwr_copy2d(const float* src, long long ld_src, float* dst, long long ld_dst, long long rows,
               int cols, int accumulate, uwr_stream_t stream);
/* column sums of a (rows, cols) matrix, deterministic two-pass; workspace >= 1024*cols floats */
int uwr_colsum(const float* x, long long ld, float* out, float* workspace, long long rows,
               int cols, uwr_stream_t stream);

/* ---- losses (LossFunction.getloss, src/Losses/losses.py:54-160) ---------------------------
 * kind: 0 "L1" (55-57), 1 "L1withColor" (58-66 + luminanceLoss.py:5-21), 2 "charbonnier"
 * (80-81,189-193), 3 "L2" (76-78), 4 mean((clamp01(p)-clamp01(t))^2) = the MSE inside torchPSNR
 * (ModelTrainer.py:17-21; grad must be NULL).  out[0] = loss, grad = dLoss/dpred (NULL to skip).
 * `batch_divisor` is the B used in the reference's "/ (B*C)" (pass the GLOBAL batch under
 * data parallelism).  workspace >= 4*1024 floats.
 */
int uwr_pixel_loss(const float* pred, const float* truth, float* out, float* grad,
                   float* workspace, int kind, int B, int C, int H, int W, int batch_divisor,
                   uwr_stream_t stream);
/* focal frequency loss (third-party focal_frequency_loss 0.3.0, ctor losses.py:48):
 * planes = B*C images of S x S (S power of two <= 1024). workspace: see *_workspace_bytes. */
size_t uwr_ffl_workspace_bytes(int planes, int S);
int uwr_ffl_loss(const float* pred, const float* truth, float* out, float* grad, float* workspace,
                 int planes, int S, uwr_stream_t stream);

/* ---- remaining "fflMix" terms and the SSIM metric on device (csrc/ssim.cu) -----------------------
 * Images are NCHW fp32, `planes` = B*C planes of H x W.
 * uwr_laplacian_l1_loss: Gradient_Loss (src/Losses/losses.py:162-181): mean |lap3x3(pred) - lap3x3(truth)| over the
 *   valid (H-2) x (W-2) region; loss[0] and grad = dloss/dpred (NULL to skip).
 * uwr_ssim_scale_fwd: one scale of pytorch_msssim (third party): separable 11-tap Gaussian
 *   (sigma 1.5), valid convolution; mean_cs[planes], mean_ss[planes] = spatial means of the cs / ssim maps;
 *   maps (planes, 3, H-10, W-10), optional: per-pixel derivatives of cs (full = 0) or ssim (full = 1) w.r.t.
 *   blur(X), blur(X^2), blur(XY) for the backward.  ssim() of ModelTrainer.torchSSIM (ModelTrainer.py:23-24) is
 *   mean(mean_ss) of one scale.
 * uwr_ssim_scale_bwd: dX = coef[plane] * d(sum of the per-pixel value)/dX (+ 0.25 * dX_coarse upsampled: the adjoint of
 *   the avg_pool2d(2) between MS-SSIM scales).
 * uwr_avgpool2_pair: avg_pool2d(2) of both images (even sides). */
size_t uwr_laplacian_l1_workspace_bytes(int planes, int H, int W);
int uwr_laplacian_l1_loss(const float* pred, const float* truth, float* loss, float* grad, float* workspace,
                          int planes, int H, int W, uwr_stream_t stream);
size_t uwr_ssim_workspace_bytes(int planes, int H, int W);
int uwr_ssim_scale_fwd(const float* X, const float* Y, float* maps, float* mean_cs, float* mean_ss,
                       float* workspace, int planes, int H, int W, float data_range, int full,
                       uwr_stream_t stream);
int uwr_ssim_scale_bwd(const float* X, const float* Y, const float* maps, const float* coef,
                       const float* dX_coarse, float* dX, int planes, int H, int W, uwr_stream_t stream);
int uwr_avgpool2_pair(const float* X, const float* Y, float* Xo, float* Yo, int planes, int H, int W,
                      uwr_stream_t stream);

/* ---- frequency-domain mixing (shared-memory FFT passes, csrc/fft.cu) ------------------------
 * Token tensors are (B, H, W, C) fp32, C contiguous; H, W (and C for *_lc_*) powers of two <= 1024.
 * uwr_dft_hw_real: y = scale * Re(FFT2 over (H, W))(x)  -- FDFP, src/model/block.py:532-556
 *                  (fftn(...).real forward: scale 1; ifftn(...).real of a real tensor: scale 1/(H*W);
 *                  the map is symmetric, so the backward of either is the same call on the cotangent).
 * uwr_dft_lc_real: y = scale * Re(FFT2 over (L = H*W tokens, C channels))(x) -- EncoderBlock,
 *                  src/model/model.py:72-88 (forward scale 1, inverse scale 1/(L*C)).
 * uwr_fft2_hw:     complex FFT2 over (H, W); in_complex 0/1, inverse 0/1 (conjugate kernel; fold the
 *                  1/(H*W) into scale); out is interleaved complex (B, H, W, C, 2) --
 *                  SpectralTransformer.UpSample, src/Models/SpectralTransformer.py:174-188.
 * workspace: uwr_dft_workspace_bytes(B, H, W, C).  x and y may not alias the workspace. */
size_t uwr_dft_workspace_bytes(int B, int H, int W, int C);
int uwr_dft_hw_real(const float* x, float* y, float* workspace, int B, int H, int W, int C,
                    float scale, uwr_stream_t stream);
int uwr_dft_lc_real(const float* x, float* y, float* workspace, int B, int H, int W, int C,
                    float scale, uwr_stream_t stream);
int uwr_fft2_hw(const float* in, float* out, float* workspace, int B, int H, int W, int C,
                int in_complex, int inverse, float scale, uwr_stream_t stream);

/* PixelShuffle(2) / PixelUnshuffle(2) on NHWC tokens (src/Models/SpectralTransformer.py:151-158,191-198;
 * src/model/block.py:107-153): in (B*H*W, 4*Cout) -> out (B*2H*2W, ld_out >= Cout) and back; H, W = the COARSE grid. */
int uwr_pixel_shuffle2(const float* in, float* out, long long ld_out, int B, int H, int W, int Cout,
                       uwr_stream_t stream);
int uwr_pixel_unshuffle2(const float* in, long long ld_in, float* out, int B, int H, int W, int Cout,
                         uwr_stream_t stream);

/* ---- thin 3x3 convolutions (stride 1, pad 1) at the 3- / 8-channel ends -----------------------
 * SpectralTransformer.embed_conv_rgb 3->16 and .output 8->3 (src/Models/SpectralTransformer.py:217,250,255,269);
 * InputProjection.proj[0] 3->8 and OutputProjection.proj[2] 8->3 (src/model/block.py:42-88).
 * A side is an NCHW image (tokens = 0) or a token matrix (B*H*W, ld) (tokens = 1).  Served: image 3 -> tokens 8|16,
 * tokens 8 -> image 3 (optionally + residual image, the `+ x` of model.py:640).  weight (Cout, Cin, 3, 3), bias may be
 * NULL.  The data gradient of tokens 8 -> image 3 is the image 3 -> tokens 8 call with flipped, transposed weights. */
typedef struct {
    const float* in;
    int in_tokens;
    long long ld_in;
    const float* weight;
    const float* bias;
    const float* residual_img;
    float* out;
    int out_tokens;
    long long ld_out;
    int B, H, W, Cin, Cout;
    int round_out; /* token output feeds a GEMM: round to TF32 in single-pass mode */
} uwr_conv_small_desc;
int uwr_conv3x3_small_fwd(const uwr_conv_small_desc* d, uwr_stream_t stream);
size_t uwr_conv3x3_small_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout);
int uwr_conv3x3_small_wgrad(const uwr_conv_small_desc* d, const float* dout, long long ld_dout, float* dweight,
                            float* dbias, float* workspace, uwr_stream_t stream);

/* ---- elementwise passes of the FFT amplitude / phase up-sampler ---------------------------------
 * SpectralTransformer.UpSample.forward, src/Models/SpectralTransformer.py:174-188.  Complex tensors are interleaved
 * (re, im) fp32 pairs; n = number of complex (split/join/cabs) or real (leaky) elements, 16-byte aligned.
 *   polar_split: mag = abs(f), pha = angle(f) (176-177); bwd: df from dmag, dpha (0 at f = 0, as torch)
 *   polar_join : z = mag*cos(pha) + i*mag*sin(pha) (181-183)
 *   cabs       : a = |z| (186)
 *   leaky_relu : LeakyReLU(slope) of amp_fuse / pha_fuse (166-169); bwd takes the saved OUTPUT
 *   even_scatter: out (B,2H,2W,C) = y (B,H,W,C) on the even pixels, bias elsewhere (post(0) = bias: the inverse
 *                 transform of the (2,2)-tiled spectrum is zero off the even pixels); even_gather is its adjoint. */
int uwr_polar_split_fwd(const float* f, float* mag, float* pha, long long n, uwr_stream_t stream);
int uwr_polar_split_bwd(const float* f, const float* dmag, const float* dpha, float* df, long long n,
                        uwr_stream_t stream);
int uwr_polar_join_fwd(const float* mag, const float* pha, float* z, long long n, uwr_stream_t stream);
int uwr_polar_join_bwd(const float* mag, const float* pha, const float* dz, float* dmag, float* dpha,
                       long long n, uwr_stream_t stream);
int uwr_cabs_fwd(const float* z, float* a, long long n, uwr_stream_t stream);
int uwr_cabs_bwd(const float* z, const float* da, float* dz, long long n, uwr_stream_t stream);
int uwr_leaky_relu_fwd(const float* x, float* y, long long n, float slope, int round_out,
                       uwr_stream_t stream); /* round_out: the result feeds a GEMM (TF32 in single-pass mode) */
int uwr_leaky_relu_bwd(const float* y, const float* dy, float* dx, long long n, float slope,
                       uwr_stream_t stream);
/* exact (erf) GELU, nn.GELU(): FDFP (src/model/block.py:541-546), Mlp.act (src/Models/AST.py:285-291) */
int uwr_gelu_fwd(const float* x, float* y, long long n, int round_out, uwr_stream_t stream);
int uwr_gelu_bwd(const float* x, const float* dy, float* dx, long long n, uwr_stream_t stream);
int uwr_even_scatter(const float* y, const float* bias, float* out, int B, int H, int W, int C,
                     uwr_stream_t stream);
int uwr_even_gather(const float* dout, float* dy, int B, int H, int W, int C, uwr_stream_t stream);

/* ---- MDTA channel attention (src/Models/SpectralTransformer.py:92-113) -----------------------
 * Token matrices are (B*L, ld) fp32, the c = C/heads channels of a head contiguous; c in {8,16,32,64},
 * heads*c <= 256, any L >= 1.  Pass pointers already offset to the first channel.
 * uwr_mdta_gram:  G[b,h,i,j] = sum_l X[b,l,h*c+i] Y[b,l,h*c+j]   (q^T k, line 100; also dA = dout^T v),
 *                 sqx[b,ch] = sum_l X[b,l,ch]^2, sqy likewise (the L2 norms of line 99; NULL to skip).
 *                 heads*c*c <= 4096.  Deterministic (per-CTA partials in `workspace`, then one reduce).
 * uwr_mdta_apply: out[b,l,h*c+i] = sum_j M[b,h,i,j] X[b,l,h*c+j]  (attn @ v, lines 101/109/113;
 *                 transpose = 1 uses M[b,h,j,i]: dv = A^T dout), optionally + diag[b,h*c+i] * Yd[b,l,h*c+i]
 *                 (the derivative of the squared norms in the backward: dq = dG k + 2 dsq_q q). */
size_t uwr_mdta_gram_workspace_bytes(int B, int L, int heads, int c);
int uwr_mdta_gram(const float* X, long long ldx, const float* Y, long long ldy, int B, int L, int heads,
                  int c, float* G, float* sqx, float* sqy, float* workspace, uwr_stream_t stream);
int uwr_mdta_apply(const float* X, long long ldx, const float* M, int transpose, const float* Yd,
                   long long ldy, const float* diag, float* out, long long ldo, int B, int L, int heads,
                   int c, uwr_stream_t stream);

/* ---- Haar "wavelet" mode (use_dwt = "Wavelet": src/model/model.py:64-88, src/model/block.py:532-552) -------
 * DWT_2D / IDWT_2D of src/model/wave_modules.py:9-181 on token tensors, forward AND the reference's own hand-written
 * backward formulas (which are not the adjoints of the forwards; csrc/wavelet.cu states them in closed form).
 * h, w = the COARSE grid; the fine grid is 2h x 2w; C % 4 == 0; idwt_bwd needs h, w % 4 == 0 and a workspace of
 * B*4*(h*w/16) floats. */
int uwr_haar_dwt_fwd(const float* in, float* out, int B, int h, int w, int C, uwr_stream_t stream);
int uwr_haar_dwt_bwd(const float* dout, float* din, int B, int h, int w, int C, uwr_stream_t stream);
int uwr_haar_idwt_fwd(const float* in, float* out, int B, int h, int w, int C, uwr_stream_t stream);
int uwr_haar_idwt_bwd(const float* dout, float* din, float* workspace, int B, int h, int w, int C,
                      uwr_stream_t stream);

/* ---- optimizer step (ModelTrainer.py:87-88,197-204): clip_grad_norm_(1.0) + Adam/AdamW -----
 * tensor tables are device arrays of pointers; `offsets` (n_tensors+1 entries, offsets[0]=0) is
 * the running element count of the virtual concatenation.
 * norm_out[0] = total L2 norm, norm_out[1] = clip coefficient min(1, max_norm/(norm+1e-6)).
 * step_dev (optional device int) overrides `step`, lr_dev (optional device float) overrides `lr`, so that a
 * captured CUDA graph can be replayed while the step counter advances and an LR scheduler (the reference uses
 * MultiStepLR([1,100,250], 0.25), ModelTrainer.py:55,129) changes the rate.
 */
int uwr_grad_norm(const float* const* grads, const long long* offsets, int n_tensors,
                  long long total_elems, float max_norm, float grad_prescale, float* norm_out,
                  float* workspace /* >= 4096 floats */, uwr_stream_t stream);
int uwr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* offsets, int n_tensors,
                  long long total_elems, const float* clip_coef /* device, may be NULL */,
                  float grad_prescale, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int decoupled, int step, const int* step_dev,
                  const float* lr_dev, uwr_stream_t stream);
int uwr_increment_i32(int* counter, uwr_stream_t stream);

/* ---- data-parallel helpers: the gradient all-reduce over NCCL / NVLink --------------------
 * Thin wrappers over the libnccl.so.2 already mapped in the process (dlopen at first use; uwr_nccl_available() = 0
 * when it is absent).  `comm` is an ncclComm_t.  uwr_nccl_allreduce_sum_f32 enqueues an in-place sum all-reduce on
 * `stream`; being a plain stream-ordered NCCL call it can be captured into a CUDA graph with the kernels around it
 * (the whole DP step = one graph, uwr/graph.py).  The reference itself is single-device (ModelTrainer.py:33-42). */
typedef struct {
    char internal[128];
} uwr_nccl_id; /* == ncclUniqueId */
int uwr_nccl_available(void);
int uwr_nccl_unique_id(uwr_nccl_id* id);                                     /* rank 0, then broadcast the 128 bytes */
int uwr_nccl_comm_init(void** comm, int nranks, const uwr_nccl_id* id, int rank); /* collective */
int uwr_nccl_comm_destroy(void* comm);
int uwr_nccl_allreduce_sum_f32(void* comm, float* buf, size_t count, uwr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UWR_B200_H */
