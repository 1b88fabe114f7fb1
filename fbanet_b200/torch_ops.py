"""``torch.library`` custom ops over the C-ABI launchers ("C-ABI ``extern "C"`` launchers wrapped as
``torch.library`` custom ops (``fbanet::warp``, ``fbanet::faf_gate_fuse``, ``fbanet::conv3x3``, ...)").

Each op is the functional form of one :mod:`fbanet_b200.ops` wrapper (it allocates its result; the model's own forward calls the
out-parameter forms directly so that concat slices and scatter stores are written in place), has a fake (meta) implementation so
shapes propagate under ``FakeTensorMode`` / ``torch.export``, and only a CUDA implementation: called with CPU tensors the
dispatcher raises ``NotImplementedError`` -- there is no CPU fallback.

==========================  =====================================================================  ==============================
op                          what                                                                   reference
==========================  =====================================================================  ==============================
``fbanet::warp``            homography warp + bilinear sampling of a burst                          homography_alignment.py:46-55
``fbanet::conv3x3``         3x3 stride-1 pad-1 conv (+bias, activation, residual), channels-last   blocks / models convs
``fbanet::linear``          per-pixel linear layer = 1x1 conv (+bias, activation, residual)        layers/linear_projection.py
``fbanet::layernorm``       LayerNorm over the channel axis                                        layers/fba_net.py:196,246
``fbanet::window_attention``windowed multi-head attention with relative-position bias + shift mask layers/window_attention.py:159-248
``fbanet::leff_mlp``        Linear -> GELU -> depthwise 3x3 -> GELU -> Linear (+ residual)         locally_enhanced_feed_forward.py
``fbanet::faf_gate_fuse``   FAF gate + F*64 -> 64 1x1 fusion + PReLU in one pass                    federated_affinity_fusion.py:67-128
``fbanet::forward``         the whole BaseModel forward of a registered model handle               models/fba_net.py:242-322
==========================  =====================================================================  ==============================
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops

__all__ = ["register_model", "OPS"]

OPS = ("warp", "conv3x3", "linear", "layernorm", "window_attention", "leff_mlp", "faf_gate_fuse", "forward")


@torch.library.custom_op("fbanet::warp", mutates_args=(), device_types="cuda")
def warp(burst: torch.Tensor, M: torch.Tensor) -> torch.Tensor:
    """``burst [B,T,C,H,W]`` fp32, ``M [B,T,3,3]`` float64 dst->src homographies (frame 0 is copied)."""
    return ops.warp_burst(burst.contiguous(), M)


@warp.register_fake
def _(burst, M):
    torch._check(burst.dim() == 5 and M.shape == (burst.shape[0], burst.shape[1], 3, 3))
    return torch.empty_like(burst)


def _conv(x, weight, bias, residual, act, k):
    N, H, W, _ = x.shape
    out = torch.empty((N, H, W, weight.shape[0]), device=x.device, dtype=x.dtype)
    return ops.conv_gemm([x], weight, out, kh=k, kw=k, pad=k // 2, bias=bias, act=act, residual=residual)


@torch.library.custom_op("fbanet::conv3x3", mutates_args=(), device_types="cuda")
def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """``x [N,H,W,Cin]`` channels-last, ``weight [Cout, 9*Cin]`` packed tap-major (``model.pack_conv_weight``), fp32 ``bias``;
    ``act``: ``FBANET_ACT_*`` of ``include/fbanet_b200.h``."""
    return _conv(x, weight, bias, residual, act, 3)


@torch.library.custom_op("fbanet::linear", mutates_args=(), device_types="cuda")
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """``x [N,H,W,Cin]``, ``weight [Cout,Cin]`` (torch ``nn.Linear`` layout)."""
    return _conv(x, weight, bias, residual, act, 1)


def _conv_fake(x, weight, bias, residual, act):
    torch._check(x.dim() == 4)
    return x.new_empty((x.shape[0], x.shape[1], x.shape[2], weight.shape[0]))


conv3x3.register_fake(_conv_fake)
linear.register_fake(_conv_fake)


@torch.library.custom_op("fbanet::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    return ops.layernorm(x.reshape(-1, x.shape[-1]), gamma, beta, eps).view(x.shape)


@layernorm.register_fake
def _(x, gamma, beta, eps):
    return torch.empty_like(x)


@torch.library.custom_op("fbanet::window_attention", mutates_args=(), device_types="cuda")
def window_attention(qkv: torch.Tensor, bias_table: torch.Tensor, heads: int, win: int, shift: int, scale: float) -> torch.Tensor:
    """``qkv [B,H,W,3C]`` (q | k | v columns) -> ``[B,H,W,C]``; ``bias_table [(2 win - 1)^2, heads]`` fp32."""
    B, H, W, C3 = qkv.shape
    return ops.window_attention(qkv.reshape(-1, C3), bias_table, B, H, W, heads, win, shift, scale).view(B, H, W, C3 // 3)


@window_attention.register_fake
def _(qkv, bias_table, heads, win, shift, scale):
    torch._check(qkv.dim() == 4 and qkv.shape[3] % 3 == 0)
    return qkv.new_empty((qkv.shape[0], qkv.shape[1], qkv.shape[2], qkv.shape[3] // 3))


@torch.library.custom_op("fbanet::leff_mlp", mutates_args=(), device_types="cuda")
def leff_mlp(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, dw_weight: torch.Tensor, dw_bias: torch.Tensor, w2: torch.Tensor,
             b2: torch.Tensor, residual: Optional[torch.Tensor], act: int) -> torch.Tensor:
    """LeFF in the layer's own parameter layout: ``w1 [4C,C]``, ``dw_weight [4C,1,3,3]``, ``w2 [C,4C]`` (fp32 or bf16 masters); dim <= 128
    runs the one-kernel MLP, larger dims the fc1 GEMM + fused depthwise/fc2 kernel."""
    N, H, W, Cc = x.shape
    Hd = w1.shape[0]
    out = torch.empty_like(x)
    dw9 = dw_weight.reshape(Hd, 9).t().float().contiguous()
    r = ops.leff_mlp(x, (w1.float() * 0.5).to(x.dtype).contiguous(), (b1.float() * 0.5).contiguous(), (dw9 * 0.5).contiguous(),
                     (dw_bias.float() * 0.5).contiguous(), w2.to(x.dtype).contiguous(), b2.float().contiguous(), out, residual, act)
    if r is None:
        h = torch.empty((N, H, W, Hd), device=x.device, dtype=x.dtype)
        ops.conv_gemm([x], w1.to(x.dtype).contiguous(), h, bias=b1.float().contiguous(), act=act)
        ops.leff_fc2(h, dw9, dw_bias.float().contiguous(), w2.to(x.dtype).contiguous(), b2.float().contiguous(), out, residual, act)
    return out


@leff_mlp.register_fake
def _(x, w1, b1, dw_weight, dw_bias, w2, b2, residual, act):
    return torch.empty_like(x)


@torch.library.custom_op("fbanet::faf_gate_fuse", mutates_args=(), device_types="cuda")
def faf_gate_fuse(feat: torch.Tensor, wsum: torch.Tensor, fuse_weight: torch.Tensor, bias: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """K2: ``feat [B,F,H,W,64]`` bf16 -> fused ``[B,H,W,64]``; ``wsum [9,64]`` = the summed score kernel (``model.py`` FAF identity),
    ``fuse_weight [64, F*64]``, PReLU slope ``alpha``."""
    B, Fr, H, W, Cc = feat.shape
    out = torch.empty((B, H, W, Cc), device=feat.device, dtype=feat.dtype)
    r = ops.faf_fuse(feat, ops.faf_fuse_score_weight(wsum), fuse_weight, bias, alpha, out)
    if r is None:
        raise RuntimeError("fbanet::faf_gate_fuse: shape not supported by the one-pass kernel")
    return out


@faf_gate_fuse.register_fake
def _(feat, wsum, fuse_weight, bias, alpha):
    return feat.new_empty((feat.shape[0], feat.shape[2], feat.shape[3], feat.shape[4]))


# whole-model op: models are registered under an integer handle (custom ops take tensors and scalars only)
_MODELS: Dict[int, "torch.nn.Module"] = {}


def register_model(model) -> int:
    """Returns the handle ``torch.ops.fbanet.forward(burst, handle)`` runs ``model`` under."""
    h = len(_MODELS)
    _MODELS[h] = model
    return h


@torch.library.custom_op("fbanet::forward", mutates_args=(), device_types="cuda")
def forward(burst: torch.Tensor, handle: int) -> torch.Tensor:
    return _MODELS[handle](burst)


@forward.register_fake
def _(burst, handle):
    torch._check(burst.dim() == 5)
    return burst.new_empty((burst.shape[0], burst.shape[2], 4 * burst.shape[3], 4 * burst.shape[4]), dtype=torch.float32)
