"""ctypes binding of ``include/fbanet_b200.h`` (the C-ABI shared library of sm_100a kernels).

There is no CPU fallback: if the library is missing, loading raises and every op fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FBANET_B200_LIB") or os.path.join(_HERE, "csrc", "libfbanet_b200.so")   # env override: A/B two builds in one process tree
ABI_VERSION = 29
MAX_SRC = 16

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2, 3, 4
STORE_NHWC, STORE_PS2, STORE_CONVT2, STORE_NCHW_BASE, STORE_NHWC_F32 = 0, 1, 2, 3, 4
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
CONVERT_U8_TO_F32, CONVERT_F32_TO_U8, CONVERT_F32_TO_F16 = 0, 1, 2

ERRORS = {-1: "bad shape", -2: "misaligned pointer/stride", -3: "unsupported dtype", -4: "CUDA launch failure", -5: "impl unsupported for this problem"}


class Src(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("row_scale", C.c_void_p), ("img_stride", C.c_int64), ("scale_img_stride", C.c_int64),
        ("C", C.c_int32), ("ld", C.c_int32),
    ]


class ConvParams(C.Structure):
    _fields_ = [
        ("src", Src * MAX_SRC),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("alpha", C.c_void_p), ("residual", C.c_void_p),
        ("out", C.c_void_p), ("base", C.c_void_p),
        ("res_img_stride", C.c_int64), ("out_img_stride", C.c_int64), ("base_img_stride", C.c_int64),
        ("dtype", C.c_int32), ("impl", C.c_int32), ("nsrc", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32), ("Cout_store", C.c_int32),
        ("act", C.c_int32), ("store_mode", C.c_int32), ("res_ld", C.c_int32), ("out_ld", C.c_int32),
        ("src_s2d", C.c_int32), ("fold_hi_lo", C.c_int32),
        ("ln_stats", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("store_f16", C.c_int32),
    ]


class WarpParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("M", C.c_void_p), ("coords", C.c_void_p),
        ("s_frame", C.c_int64), ("s_y", C.c_int64), ("s_x", C.c_int64), ("s_c", C.c_int64),
        ("d_frame", C.c_int64), ("d_y", C.c_int64), ("d_x", C.c_int64), ("d_c", C.c_int64),
        ("frames", C.c_int32), ("frames_per_burst", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("_pad", C.c_int32),
    ]


class ToNhwcParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("dtype", C.c_int32),
        ("frames", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cp", C.c_int32),
        ("im2col3x3", C.c_int32), ("_pad", C.c_int32),
    ]


class S2dParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("img_stride", C.c_int64), ("dtype", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("ld", C.c_int32),
    ]


class HeadConvParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("dtype", C.c_int32),
        ("frames", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cout", C.c_int32),
        ("frames_per_burst", C.c_int32), ("M", C.c_void_p),
    ]


class AssembleParams(C.Structure):
    _fields_ = [
        ("sr", C.c_void_p), ("base", C.c_void_p), ("out", C.c_void_p), ("base_img_stride", C.c_int64), ("dtype", C.c_int32),
        ("N", C.c_int32), ("C", C.c_int32), ("Cp", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("lo_offset", C.c_int32), ("_pad", C.c_int32),
    ]


class ConvertIoParams(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int64), ("mode", C.c_int32), ("_pad", C.c_int32)]


class LayerNormParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("rows", C.c_int64),
        ("C", C.c_int32), ("x_ld", C.c_int32), ("y_ld", C.c_int32), ("dtype", C.c_int32), ("eps", C.c_float), ("_pad", C.c_int32),
        ("stats", C.c_void_p),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("out", C.c_void_p), ("bias_table", C.c_void_p), ("dtype", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("heads", C.c_int32), ("win", C.c_int32), ("shift", C.c_int32),
        ("qkv_ld", C.c_int32), ("out_ld", C.c_int32), ("scale", C.c_float), ("impl", C.c_int32), ("bias_expanded", C.c_void_p),
        ("q_prescaled", C.c_int32), ("_pad", C.c_int32), ("bias_wrap", C.c_void_p),
    ]


class DwconvParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("dtype", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("act", C.c_int32),
    ]


class FafGateParams(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("gate", C.c_void_p), ("wsum", C.c_void_p), ("gated", C.c_void_p), ("dtype", C.c_int32),
        ("B", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("_pad", C.c_int32),
        ("score", C.c_void_p),
    ]


class FafFuseParams(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("score_weight", C.c_void_p), ("fuse_weight", C.c_void_p), ("bias", C.c_void_p), ("alpha", C.c_void_p),
        ("gate", C.c_void_p), ("out", C.c_void_p), ("out_img_stride", C.c_int64), ("out_ld", C.c_int32),
        ("B", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("_pad", C.c_int32 * 2),
    ]


class LeffMlpParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w1", C.c_void_p), ("bias1", C.c_void_p), ("dw_weight", C.c_void_p), ("dw_bias", C.c_void_p), ("w2", C.c_void_p),
        ("bias2", C.c_void_p), ("residual", C.c_void_p), ("out", C.c_void_p),
        ("x_img_stride", C.c_int64), ("res_img_stride", C.c_int64), ("out_img_stride", C.c_int64),
        ("x_ld", C.c_int32), ("res_ld", C.c_int32), ("out_ld", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("Hd", C.c_int32), ("act", C.c_int32), ("w2_f16", C.c_int32),
    ]


class LeffFc2Params(C.Structure):
    _fields_ = [
        ("h1", C.c_void_p), ("dw_weight", C.c_void_p), ("dw_bias", C.c_void_p), ("w2", C.c_void_p), ("bias2", C.c_void_p),
        ("residual", C.c_void_p), ("out", C.c_void_p), ("res_img_stride", C.c_int64), ("out_img_stride", C.c_int64),
        ("res_ld", C.c_int32), ("out_ld", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("Hd", C.c_int32), ("act", C.c_int32),
        ("f16", C.c_int32), ("_pad", C.c_int32),
    ]


class TileParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p),
        ("T", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("psize", C.c_int32), ("overlap", C.c_int32), ("tile_begin", C.c_int32), ("tile_end", C.c_int32),
        ("scale", C.c_int32), ("_pad", C.c_int32),
    ]


MAX_BANDS = 8


class TileBandParams(C.Structure):
    _fields_ = [
        ("band", C.c_void_p * MAX_BANDS), ("tiles", C.c_void_p), ("row0", C.c_int32 * (MAX_BANDS + 1)), ("nbands", C.c_int32),
        ("T", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("psize", C.c_int32), ("overlap", C.c_int32), ("tile_begin", C.c_int32), ("tile_end", C.c_int32),
        ("scale", C.c_int32), ("_pad", C.c_int32),
    ]


class FlowWarpParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("flow", C.c_void_p),
        ("s_frame", C.c_int64), ("s_y", C.c_int64), ("s_x", C.c_int64), ("s_c", C.c_int64),
        ("d_frame", C.c_int64), ("d_y", C.c_int64), ("d_x", C.c_int64), ("d_c", C.c_int64),
        ("frames", C.c_int32), ("frames_per_burst", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("_pad", C.c_int32),
    ]


class EccPrepareParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("planes", C.c_void_p),
        ("s_frame", C.c_int64), ("s_y", C.c_int64), ("s_x", C.c_int64), ("s_c", C.c_int64),
        ("gray_weight", C.c_float * 4), ("frames", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
    ]


class EccParams(C.Structure):
    _fields_ = [
        ("planes", C.c_void_p), ("warp", C.c_void_p), ("rho", C.c_void_p), ("iters_done", C.c_void_p), ("eps", C.c_double),
        ("frames", C.c_int32), ("frames_per_burst", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("max_iters", C.c_int32), ("_pad", C.c_int32),
    ]


class TrainLossParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("grad", C.c_void_p), ("partial", C.c_void_p), ("loss", C.c_void_p),
        ("eps", C.c_float), ("gw_weight", C.c_float), ("inv_n", C.c_float), ("planes", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("clamp_restored", C.c_int32), ("_pad", C.c_int32),
    ]


class AdamParams(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("n", C.c_int64),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("step_size", C.c_float), ("bias2_sqrt", C.c_float), ("grad_scale", C.c_float), ("one_minus_beta1", C.c_float), ("one_minus_beta2", C.c_float),
        ("decoupled", C.c_int32), ("_pad", C.c_int32),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p), ("db", C.c_void_p), ("partial", C.c_void_p),
        ("x_img_stride", C.c_int64), ("dy_img_stride", C.c_int64), ("x_ld", C.c_int32), ("dy_ld", C.c_int32),
        ("dtype", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("Cout", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("splits", C.c_int32),
        ("accumulate", C.c_int32),
    ]


class LayerNormBwdParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("gamma", C.c_void_p), ("dx", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("partial", C.c_void_p), ("rows", C.c_int64), ("eps", C.c_float), ("dtype", C.c_int32), ("C", C.c_int32), ("accumulate", C.c_int32),
    ]


class ActBwdParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p), ("alpha", C.c_void_p), ("dalpha", C.c_void_p), ("partial", C.c_void_p),
        ("n", C.c_int64), ("dtype", C.c_int32), ("act", C.c_int32), ("accumulate", C.c_int32), ("_pad", C.c_int32),
    ]


class ActFwdParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("alpha", C.c_void_p), ("n", C.c_int64), ("dtype", C.c_int32), ("act", C.c_int32)]


class DwconvBwdParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("weight", C.c_void_p), ("dx", C.c_void_p), ("dw", C.c_void_p), ("db", C.c_void_p),
        ("partial", C.c_void_p), ("dtype", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("accumulate", C.c_int32),
    ]


class AttnBwdParams(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("dout", C.c_void_p), ("dqkv", C.c_void_p), ("bias_table", C.c_void_p), ("dbias", C.c_void_p),
        ("partial", C.c_void_p), ("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("heads", C.c_int32), ("win", C.c_int32), ("shift", C.c_int32), ("qkv_ld", C.c_int32), ("dout_ld", C.c_int32),
        ("dqkv_ld", C.c_int32), ("scale", C.c_float), ("accumulate", C.c_int32), ("_pad", C.c_int32),
    ]


class FafGateBwdParams(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("dgated", C.c_void_p), ("gate", C.c_void_p), ("score", C.c_void_p), ("wsum", C.c_void_p),
        ("dfeat", C.c_void_p), ("dscore", C.c_void_p), ("dwsum", C.c_void_p), ("partial", C.c_void_p),
        ("dtype", C.c_int32), ("B", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("accumulate", C.c_int32), ("_pad", C.c_int32),
    ]


class DropPathParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("skip", C.c_void_p), ("out", C.c_void_p), ("scale", C.c_void_p), ("per_burst", C.c_int64),
        ("dtype", C.c_int32), ("B", C.c_int32),
    ]


STRUCTS = {
    "fbanet_src": Src, "fbanet_conv_params": ConvParams, "fbanet_warp_params": WarpParams,
    "fbanet_to_nhwc_params": ToNhwcParams, "fbanet_s2d_params": S2dParams, "fbanet_head_conv_params": HeadConvParams, "fbanet_assemble_params": AssembleParams, "fbanet_convert_io_params": ConvertIoParams, "fbanet_layernorm_params": LayerNormParams, "fbanet_attn_params": AttnParams,
    "fbanet_dwconv_params": DwconvParams, "fbanet_faf_gate_params": FafGateParams, "fbanet_faf_fuse_params": FafFuseParams, "fbanet_leff_fc2_params": LeffFc2Params, "fbanet_leff_mlp_params": LeffMlpParams, "fbanet_tile_params": TileParams,
    "fbanet_tile_band_params": TileBandParams, "fbanet_flow_warp_params": FlowWarpParams,
    "fbanet_ecc_prepare_params": EccPrepareParams, "fbanet_ecc_params": EccParams, "fbanet_train_loss_params": TrainLossParams,
    "fbanet_adam_params": AdamParams, "fbanet_wgrad_params": WgradParams, "fbanet_layernorm_bwd_params": LayerNormBwdParams,
    "fbanet_act_bwd_params": ActBwdParams, "fbanet_act_fwd_params": ActFwdParams, "fbanet_dwconv_bwd_params": DwconvBwdParams, "fbanet_attn_bwd_params": AttnBwdParams,
    "fbanet_faf_gate_bwd_params": FafGateBwdParams, "fbanet_drop_path_params": DropPathParams,
}

# every symbol include/fbanet_b200.h declares
OPS = {
    "fbanet_warp_sm100": WarpParams, "fbanet_to_nhwc_sm100": ToNhwcParams, "fbanet_space_to_depth_sm100": S2dParams, "fbanet_head_conv_sm100": HeadConvParams, "fbanet_assemble_sm100": AssembleParams, "fbanet_convert_io_sm100": ConvertIoParams, "fbanet_conv_gemm_sm100": ConvParams,
    "fbanet_layernorm_sm100": LayerNormParams, "fbanet_window_attention_sm100": AttnParams, "fbanet_dwconv3x3_sm100": DwconvParams,
    "fbanet_faf_gate_sm100": FafGateParams, "fbanet_faf_fuse_sm100": FafFuseParams, "fbanet_leff_fc2_sm100": LeffFc2Params, "fbanet_leff_mlp_sm100": LeffMlpParams, "fbanet_tile_divide_sm100": TileParams, "fbanet_tile_merge_sm100": TileParams,
    "fbanet_tile_divide_banded_sm100": TileBandParams, "fbanet_tile_merge_banded_sm100": TileBandParams, "fbanet_flow_warp_sm100": FlowWarpParams,
    "fbanet_ecc_prepare_sm100": EccPrepareParams, "fbanet_ecc_homography_sm100": EccParams, "fbanet_train_loss_sm100": TrainLossParams,
    "fbanet_adam_step_sm100": AdamParams, "fbanet_wgrad_sm100": WgradParams, "fbanet_layernorm_bwd_sm100": LayerNormBwdParams,
    "fbanet_act_bwd_sm100": ActBwdParams, "fbanet_act_fwd_sm100": ActFwdParams, "fbanet_dwconv3x3_bwd_sm100": DwconvBwdParams,
    "fbanet_window_attention_bwd_sm100": AttnBwdParams, "fbanet_faf_gate_bwd_sm100": FafGateBwdParams,
    "fbanet_drop_path_add_sm100": DropPathParams,
}
MISC_SYMBOLS = ["fbanet_abi_version", "fbanet_abi_sizeof", "fbanet_last_cuda_error", "fbanet_conv_gemm_tcgen05_supported", "fbanet_leff_fc2_supported", "fbanet_leff_mlp_supported", "fbanet_faf_fuse_supported", "fbanet_window_attention_tcgen05_supported",
                "fbanet_train_loss_workspace_doubles", "fbanet_layernorm_bwd_blocks", "fbanet_act_bwd_blocks",
                "fbanet_dwconv_bwd_blocks", "fbanet_faf_gate_bwd_blocks", "fbanet_attn_bwd_partial_floats"]

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError when it is absent or its ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fbanet_b200: CUDA library {LIB_PATH} is missing -- run `python -m fbanet_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.fbanet_abi_version.restype = C.c_int
    lib.fbanet_abi_sizeof.restype = C.c_int
    lib.fbanet_abi_sizeof.argtypes = [C.c_char_p]
    lib.fbanet_last_cuda_error.restype = C.c_char_p
    lib.fbanet_conv_gemm_tcgen05_supported.restype = C.c_int
    lib.fbanet_conv_gemm_tcgen05_supported.argtypes = [C.POINTER(ConvParams)]
    lib.fbanet_leff_fc2_supported.restype = C.c_int
    lib.fbanet_leff_fc2_supported.argtypes = [C.POINTER(LeffFc2Params)]
    lib.fbanet_leff_mlp_supported.restype = C.c_int
    lib.fbanet_leff_mlp_supported.argtypes = [C.POINTER(LeffMlpParams)]
    lib.fbanet_window_attention_tcgen05_supported.restype = C.c_int
    lib.fbanet_window_attention_tcgen05_supported.argtypes = [C.POINTER(AttnParams)]
    lib.fbanet_faf_fuse_supported.restype = C.c_int
    lib.fbanet_faf_fuse_supported.argtypes = [C.POINTER(FafFuseParams)]
    lib.fbanet_train_loss_workspace_doubles.restype = C.c_int64
    lib.fbanet_train_loss_workspace_doubles.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    for fn in (lib.fbanet_layernorm_bwd_blocks, lib.fbanet_act_bwd_blocks, lib.fbanet_dwconv_bwd_blocks, lib.fbanet_faf_gate_bwd_blocks):
        fn.restype = C.c_int32
        fn.argtypes = [C.c_int64]
    lib.fbanet_attn_bwd_partial_floats.restype = C.c_int64
    lib.fbanet_attn_bwd_partial_floats.argtypes = [C.c_int32] * 5
    if lib.fbanet_abi_version() != ABI_VERSION:
        raise RuntimeError(f"fbanet_b200: ABI mismatch (library {lib.fbanet_abi_version()}, binding {ABI_VERSION}); rebuild")
    for name, st in STRUCTS.items():
        n = lib.fbanet_abi_sizeof(name.encode())
        if n != C.sizeof(st):
            raise RuntimeError(f"fbanet_b200: struct {name} is {n} bytes in the library but {C.sizeof(st)} in the binding")
    for name, st in OPS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(st), C.c_void_p]
    _lib = lib
    return lib


def call(name: str, params, stream: int) -> None:
    lib = load()
    rc = getattr(lib, name)(C.byref(params), C.c_void_p(stream))
    if rc != 0:
        msg = ERRORS.get(rc, f"error {rc}")
        if rc == -4:
            msg += ": " + lib.fbanet_last_cuda_error().decode()
        raise RuntimeError(f"{name} failed: {msg}")
