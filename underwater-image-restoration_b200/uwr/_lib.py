"""ctypes binding of libuwr_b200.so (the C ABI declared in include/uwr_b200.h).

There is deliberately NO fallback: if the CUDA library is missing the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "csrc", "libuwr_b200.so")
LIB_PATH = os.path.normpath(os.environ.get("UWR_B200_LIB", LIB_PATH))  # override: instrumented debug builds only

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"uwr: native library {LIB_PATH} is missing. Build it with "
        "`python underwater-image-restoration_b200/build.py` (nvcc, sm_100a). "
        "There is no CPU/PyTorch fallback for the hot path.")

lib = C.CDLL(LIB_PATH)

c_fp = C.c_void_p      # device float*
c_ll = C.c_longlong
c_int = C.c_int
c_f = C.c_float
c_sz = C.c_size_t
c_stream = C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", c_fp), ("lda", c_ll), ("a_km", c_int),
        ("B", c_fp), ("ldb", c_ll), ("b_nk", c_int),
        ("B2", c_fp), ("n_split", c_int),
        ("C", c_fp), ("ldc", c_ll),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("bias", c_fp), ("bias2", c_fp),
        ("epilogue", c_int),
        ("R", c_fp), ("ldr", c_ll),
        ("rowscale", c_fp), ("rows_per_group", c_int),
        ("colsum", c_fp),
        ("workspace", c_fp), ("workspace_bytes", c_sz),
        ("round_out", c_int),
        ("c_half", c_int), ("r_half", c_int),
    ]


class ConvSmallDesc(C.Structure):
    _fields_ = [
        ("inp", c_fp), ("in_tokens", c_int), ("ld_in", c_ll),
        ("weight", c_fp), ("bias", c_fp), ("residual_img", c_fp),
        ("out", c_fp), ("out_tokens", c_int), ("ld_out", c_ll),
        ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
        ("round_out", c_int),
    ]


class NcclId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", c_fp), ("ld_q", c_ll), ("q_off", c_int),
        ("kv", c_fp), ("ld_kv", c_ll), ("k_off", c_int), ("v_off", c_int),
        ("bias_table", c_fp), ("w_param", c_fp),
        ("B", c_int), ("H", c_int), ("W", c_int), ("heads", c_int), ("head_dim", c_int),
        ("shift", c_int), ("scale", c_f), ("operands_rounded", c_int),
        ("dq_colsum", c_fp), ("dkv_colsum", c_fp),
    ]


class ConvGemmDesc(C.Structure):
    _fields_ = [
        ("mode", c_int), ("x", c_fp), ("ld_x", c_ll),
        ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int),
        ("kh", c_int), ("kw", c_int), ("stride", c_int), ("pad", c_int), ("Cout", c_int),
        ("w", c_fp), ("bias", c_fp), ("y", c_fp), ("ld_y", c_ll), ("round_out", c_int),
        ("dy", c_fp), ("ld_dy", c_ll), ("dw", c_fp), ("workspace", c_fp), ("workspace_bytes", c_sz),
    ]


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


SIGNATURES = {
    "uwr_last_error": (C.c_char_p, []),
    "uwr_abi_version": (c_int, []),
    "uwr_device_sm_count": (c_int, []),
    "uwr_launch_count": (C.c_ulonglong, []),
    "uwr_set_gemm_precision": (c_int, [c_int]),
    "uwr_get_gemm_precision": (c_int, []),
    "uwr_gemm_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "uwr_gemm_tf32": (c_int, [C.POINTER(GemmDesc), c_stream]),
    "uwr_gemm_tcgen05": (c_int, [C.POINTER(GemmDesc), c_stream]),
    "uwr_set_gemm_cluster": (c_int, [c_int]),
    "uwr_set_pdl": (c_int, [c_int]),
    "uwr_gemm_tcgen05_supported": (c_int, [C.POINTER(GemmDesc)]),
    "uwr_convgemm_tcgen05_supported": (c_int, [C.POINTER(ConvGemmDesc)]),
    "uwr_convgemm_tcgen05_workspace_bytes": (c_sz, [C.POINTER(ConvGemmDesc)]),
    "uwr_convgemm_tcgen05": (c_int, [C.POINTER(ConvGemmDesc), c_stream]),
    "uwr_gemm_tcgen05_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "uwr_round_tf32_tensors": (c_int, [c_fp, c_fp, c_fp, c_int, c_ll, c_int, c_stream]),
    "uwr_scale_round": (c_int, [c_fp, c_ll, c_fp, c_ll, c_int, c_fp, c_int, c_int, c_stream]),
    "uwr_scale_round_colsum": (c_int, [c_fp, c_ll, c_fp, c_ll, c_int, c_fp, c_int, c_int, c_fp, c_fp, c_stream]),
    "uwr_layernorm_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_ll, c_int, c_f, c_stream]),
    "uwr_layernorm_bwd_workspace_bytes": (c_sz, [c_ll, c_int]),
    "uwr_layernorm_bwd": (c_int, [c_fp] * 10 + [c_ll, c_int, c_stream]),
    "uwr_layernorm_bwd_ds_supported": (c_int, [c_ll, c_int]),
    "uwr_layernorm_bwd_ds_workspace_bytes": (c_sz, [c_ll, c_int]),
    "uwr_layernorm_bwd_ds": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp,
                                     c_fp, c_ll, c_int, c_stream]),
    "uwr_window_attn_fwd": (c_int, [C.POINTER(AttnDesc), c_fp, c_ll, c_stream]),
    "uwr_window_attn_bwd_workspace_bytes": (c_sz, [C.POINTER(AttnDesc)]),
    "uwr_window_attn_bwd": (c_int, [C.POINTER(AttnDesc), c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp, c_stream]),
    "uwr_dwconv_gelu_fwd": (c_int, [c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_stream]),
    "uwr_set_attn_tcgen05": (c_int, [c_int]),
    "uwr_dwconv_half_supported": (c_int, [c_int, c_int, c_int]),
    "uwr_dwconv_gelu_fwd_half": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_dwconv_gelu_bwd_half": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int,
                                         c_stream]),
    "uwr_gelu_mul_fwd": (c_int, [c_fp, c_ll, c_fp, c_ll, c_int, c_stream]),
    "uwr_gelu_mul_bwd": (c_int, [c_fp, c_fp, c_ll, c_fp, c_ll, c_int, c_stream]),
    "uwr_dwconv_gelu_bwd_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "uwr_dwconv_gelu_bwd": (c_int, [c_fp, c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                    c_int, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_gelu_gate_bwd": (c_int, [c_fp, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_int, c_int, c_stream]),
    "uwr_input_proj_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_f, c_stream]),
    "uwr_input_proj_bwd_workspace_bytes": (c_sz, [c_int] * 5),
    "uwr_input_proj_bwd": (c_int, [c_fp] * 6 + [c_int] * 5 + [c_f, c_stream]),
    "uwr_output_proj_fwd": (c_int, [c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_output_proj_bwd_workspace_bytes": (c_sz, [c_int] * 4),
    "uwr_output_proj_bwd": (c_int, [c_fp, c_fp, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp,
                                    c_int, c_int, c_int, c_int, c_stream]),
    "uwr_im2col_4x4s2": (c_int, [c_fp, c_ll, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_col2im_4x4s2": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_im2col_3x3": (c_int, [c_fp, c_ll, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_col2im_3x3": (c_int, [c_fp, c_fp, c_ll, c_int, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_pixel_scatter_2x2": (c_int, [c_fp, c_fp, c_fp, c_ll, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_pixel_gather_2x2": (c_int, [c_fp, c_ll, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_copy2d": (c_int, [c_fp, c_ll, c_fp, c_ll, c_ll, c_int, c_int, c_stream]),
    "uwr_colsum": (c_int, [c_fp, c_ll, c_fp, c_fp, c_ll, c_int, c_stream]),
    "uwr_pixel_loss": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_laplacian_l1_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "uwr_laplacian_l1_loss": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_stream]),
    "uwr_ssim_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "uwr_ssim_scale_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_f, c_int, c_stream]),
    "uwr_ssim_scale_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_stream]),
    "uwr_avgpool2_pair": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_stream]),
    "uwr_ffl_workspace_bytes": (c_sz, [c_int, c_int]),
    "uwr_ffl_loss": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_stream]),
    "uwr_dft_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "uwr_dft_hw_real": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_f, c_stream]),
    "uwr_dft_lc_real": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_f, c_stream]),
    "uwr_fft2_hw": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_f, c_stream]),
    "uwr_pixel_shuffle2": (c_int, [c_fp, c_fp, c_ll, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_pixel_unshuffle2": (c_int, [c_fp, c_ll, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_conv3x3_small_fwd": (c_int, [C.POINTER(ConvSmallDesc), c_stream]),
    "uwr_conv3x3_small_wgrad_workspace_bytes": (c_sz, [c_int] * 5),
    "uwr_conv3x3_small_wgrad": (c_int, [C.POINTER(ConvSmallDesc), c_fp, c_ll, c_fp, c_fp, c_fp, c_stream]),
    "uwr_polar_split_fwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_polar_split_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_polar_join_fwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_polar_join_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_cabs_fwd": (c_int, [c_fp, c_fp, c_ll, c_stream]),
    "uwr_cabs_bwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_leaky_relu_fwd": (c_int, [c_fp, c_fp, c_ll, c_f, c_int, c_stream]),
    "uwr_gelu_fwd": (c_int, [c_fp, c_fp, c_ll, c_int, c_stream]),
    "uwr_gelu_bwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_stream]),
    "uwr_leaky_relu_bwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_f, c_stream]),
    "uwr_even_scatter": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_even_gather": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_mdta_gram_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int]),
    "uwr_mdta_gram": (c_int, [c_fp, c_ll, c_fp, c_ll, c_int, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_fp, c_stream]),
    "uwr_mdta_apply": (c_int, [c_fp, c_ll, c_fp, c_int, c_fp, c_ll, c_fp, c_fp, c_ll, c_int, c_int, c_int, c_int,
                               c_stream]),
    "uwr_haar_dwt_fwd": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_haar_dwt_bwd": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_haar_idwt_fwd": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_haar_idwt_bwd": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_stream]),
    "uwr_nccl_available": (c_int, []),
    "uwr_nccl_unique_id": (c_int, [C.POINTER(NcclId)]),
    "uwr_nccl_comm_init": (c_int, [C.POINTER(C.c_void_p), c_int, C.POINTER(NcclId), c_int]),
    "uwr_nccl_comm_destroy": (c_int, [C.c_void_p]),
    "uwr_nccl_allreduce_sum_f32": (c_int, [C.c_void_p, c_fp, c_sz, c_stream]),
    "uwr_grad_norm": (c_int, [c_fp, c_fp, c_int, c_ll, c_f, c_f, c_fp, c_fp, c_stream]),
    "uwr_adam_step": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_ll, c_fp, c_f, c_f, c_f, c_f, c_f, c_f,
                              c_int, c_int, c_fp, c_fp, c_stream]),
    "uwr_increment_i32": (c_int, [c_fp, c_stream]),
}

fn = {name: _sig(name, *sig) for name, sig in SIGNATURES.items()}


class UwrError(RuntimeError):
    pass


def check(rc, who):
    if rc != 0:
        msg = fn["uwr_last_error"]()
        raise UwrError(f"{who} failed ({rc}): {msg.decode() if msg else ''}")
