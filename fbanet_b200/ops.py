"""Thin torch-tensor wrappers over the C-ABI kernels (``include/fbanet_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; every op is a hand-written sm_100a
kernel launched through ctypes on ``torch.cuda.current_stream()``.  No op has a CPU or eager fallback.

Channels-last views: a tensor ``[N,H,W,C]`` with ``stride(-1) == 1`` and ``stride(1) == W*stride(2)``
(channel slices of wider buffers are fine -- that is how concat-free skip connections are wired).
"""
from __future__ import annotations

import math

import ctypes as C
from typing import Tuple, Optional, Sequence

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}

# launch counter: bench.py reports how many of OUR kernels ran in the timed region
LAUNCHES = 0
LAST_ATTENTION_ON_TCGEN05 = False   # whether the last window_attention call took the tcgen05 / TMEM kernel (tests)
# when a list, conv_gemm appends (start_event, end_event, algorithmic_flops) per launch (bench roofline)
_PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# when a list, every launch appends (op name, start_event, end_event, tag) -- per-kernel breakdowns
_OPPROF = None


def _call(name, p, tag="", nbytes=0):
    """Launch one C-ABI op on the current stream.  ``nbytes``: ALGORITHMIC HBM bytes of the launch (tensors that must be read
    once + written once), used by the roofline accounting of ``profile_ops``."""
    global LAUNCHES
    LAUNCHES += 1
    if _OPPROF is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call(name, p, _stream())
        e1.record()
        _OPPROF.append((name, e0, e1, tag, nbytes))
        return
    L.call(name, p, _stream())


def profile_ops(fn, stream, by_tag: bool = True) -> dict:
    """Run ``fn`` once with CUDA events around every kernel launch; {op name[:tag]: (total ms, launches, algorithmic bytes)}."""
    global _OPPROF
    _OPPROF = []
    try:
        fn()
        stream.synchronize()
        out = {}
        for name, a, b, tag, nbytes in _OPPROF:
            k = name + (":" + tag if (tag and by_tag) else "")
            ms, n, by = out.get(k, (0.0, 0, 0))
            out[k] = (ms + a.elapsed_time(b), n + 1, by + nbytes)
    finally:
        _OPPROF = None
    return out


def _cl(t: torch.Tensor):
    """(ptr, C, ld, img_stride) of a channels-last view ``[N,H,W,C]``."""
    assert t.is_cuda and t.dim() == 4, "expected a CUDA [N,H,W,C] view"
    N, H, W, Cc = t.shape
    sN, sH, sW, sC = t.stride()
    assert sC == 1 and (H == 1 or sH == W * sW), f"not a channels-last view: shape {tuple(t.shape)} stride {t.stride()}"
    if N == 1:
        sN = H * W * sW
    return t.data_ptr(), Cc, sW, sN


def conv_gemm(
    srcs: Sequence[torch.Tensor],
    weight: torch.Tensor,
    out: torch.Tensor,
    *,
    kh: int = 1,
    kw: int = 1,
    stride: int = 1,
    pad: int = 0,
    bias: Optional[torch.Tensor] = None,
    act: int = L.ACT_NONE,
    alpha: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    store_mode: int = L.STORE_NHWC,
    row_scales: Optional[Sequence[Optional[torch.Tensor]]] = None,
    base: Optional[torch.Tensor] = None,
    cout_store: Optional[int] = None,
    impl: int = L.IMPL_AUTO,
    alg_cin: Optional[int] = None,
    src_s2d: bool = False,
    ln_stats: Optional[torch.Tensor] = None,
    fold_hi_lo: bool = False,
    ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
    ln_eps: float = 1e-5,
) -> torch.Tensor:
    """out = act(conv(concat(srcs)) + bias) + residual.  ``weight`` is packed ``[Cout, kh*kw*sum(C)]``.

    ``ln_stats`` ``[rows, 2]`` fp32 (mean, rstd per input row, :func:`row_stats`): LayerNorm folded into a 1x1 GEMM -- ``weight``
    must be the row-centred ``W * gamma`` and ``bias`` ``W @ beta + b`` (:func:`fold_layernorm`); the result equals
    ``W @ LN(x) + b``.

    ``ln = (gamma, beta)`` (fp32 ``[C]``): LayerNorm applied to the single source INSIDE the 1x1 GEMM -- the raw rows are normalised in
    shared memory with the LayerNorm kernel's own arithmetic before the tensor core reads them; ``weight`` / ``bias`` are the layer's
    own.  Bit-identical to :func:`layernorm` + GEMM.  Raises ``RuntimeError("... impl unsupported ...")`` when the shape does not fit
    (callers then run the two kernels)."""
    p = L.ConvParams()
    dt = srcs[0].dtype
    p.dtype = _DT[dt]
    p.impl = impl
    p.nsrc = len(srcs)
    assert 1 <= len(srcs) <= L.MAX_SRC
    N, H, W, _ = srcs[0].shape
    p.src_s2d = 1 if src_s2d else 0
    ctot = 0
    for i, s in enumerate(srcs):
        assert s.dtype == dt and s.shape[:3] == (N, H, W)
        ptr, Cc, ld, istr = _cl(s)
        p.src[i].ptr, p.src[i].C, p.src[i].ld, p.src[i].img_stride = ptr, Cc, ld, istr
        ctot += Cc
        rs = row_scales[i] if row_scales is not None else None
        if rs is not None:  # [N,H,W] fp32 view, contiguous inside an image
            assert rs.dtype == torch.float32 and rs.shape == (N, H, W) and rs.stride(2) == 1 and rs.stride(1) == W
            p.src[i].row_scale = rs.data_ptr()
            p.src[i].scale_img_stride = rs.stride(0) if N > 1 else H * W
    if src_s2d:  # sources are [N,H/2,W/2,4C] space-to-depth views of the logical [N,H,W,C] inputs
        ctot //= 4
        H, W = 2 * H, 2 * W
    cout = weight.shape[0]
    assert weight.dtype == dt and weight.is_contiguous() and weight.shape[1] == kh * kw * ctot, (weight.shape, kh, kw, ctot)
    p.weight = weight.data_ptr()
    p.N, p.H, p.W = N, H, W
    p.KH, p.KW, p.stride, p.pad = kh, kw, stride, pad
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    p.Ho, p.Wo, p.Cout = Ho, Wo, cout
    p.Cout_store = cout if cout_store is None else cout_store
    p.act, p.store_mode = act, store_mode
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout
        p.bias = bias.data_ptr()
    if alpha is not None:
        assert alpha.dtype == torch.float32
        p.alpha = alpha.data_ptr()
    if ln is not None:
        g_, b_ = ln
        assert ln_stats is None and len(srcs) == 1 and kh == 1 and kw == 1
        assert g_.dtype == torch.float32 and b_.dtype == torch.float32 and g_.numel() == ctot and b_.numel() == ctot and g_.is_contiguous() and b_.is_contiguous()
        p.ln_gamma, p.ln_beta, p.ln_eps = g_.data_ptr(), b_.data_ptr(), ln_eps
    if ln_stats is not None:
        assert ln_stats.dtype == torch.float32 and ln_stats.is_contiguous() and ln_stats.shape == (N * Ho * Wo, 2) and kh == 1 and kw == 1
        p.ln_stats = ln_stats.data_ptr()
    if residual is not None:
        assert residual.dtype == dt and residual.shape == (N, Ho, Wo, cout)
        rp, _, rld, ris = _cl(residual)
        p.residual, p.res_ld, p.res_img_stride = rp, rld, ris
    if store_mode == L.STORE_NCHW_BASE:
        assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (N, p.Cout_store, Ho, Wo)
        assert base is not None and base.dtype == torch.float32 and base.shape == (N, p.Cout_store, Ho // 4, Wo // 4)
        assert base.stride(3) == 1 and base.stride(2) == Wo // 4 and base.stride(1) == (Ho // 4) * (Wo // 4)
        p.out, p.out_img_stride = out.data_ptr(), p.Cout_store * Ho * Wo
        p.base, p.base_img_stride = base.data_ptr(), (base.stride(0) if N > 1 else 0)
    elif store_mode == L.STORE_NHWC_F32:
        if fold_hi_lo:   # hi / lo weight rows summed by the kernel: Cout_store / 2 outputs (3 are stored as 4 with a zero column)
            p.fold_hi_lo = 1
            assert out.dtype == torch.float32 and out.shape == (N, Ho, Wo, 4 if p.Cout_store // 2 == 3 else p.Cout_store // 2)
        else:
            assert out.dtype == torch.float32 and out.shape == (N, Ho, Wo, p.Cout_store)
        op, _, old, ois = _cl(out)
        p.out, p.out_ld, p.out_img_stride = op, old, ois
    else:
        if out.dtype == torch.float16 and dt == torch.bfloat16:   # fp16 store of the staged epilogue (the dim-256 LeFF's hidden map)
            assert store_mode == L.STORE_NHWC and residual is None
            p.store_f16 = 1
        else:
            assert out.dtype == dt
        if store_mode == L.STORE_NHWC:
            assert out.shape == (N, Ho, Wo, p.Cout_store)
        elif store_mode == L.STORE_PS2:
            assert out.shape == (N, 2 * Ho, 2 * Wo, cout // 4)
        else:
            assert out.shape == (N, 2 * Ho, 2 * Wo, cout // 4)
        op, _, old, ois = _cl(out)
        p.out, p.out_ld, p.out_img_stride = op, old, ois
    esz = srcs[0].element_size()
    nbytes = N * H * W * ctot * esz + out.numel() * out.element_size() + (residual.numel() * esz if residual is not None else 0) + weight.numel() * esz
    tag = f"k{kh}s{stride} {ctot}->{cout} @{Ho}x{Wo} act{act}{' res' if residual is not None else ''} st{store_mode}{' ln' if ln is not None else ''}"
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _call("fbanet_conv_gemm_sm100", p)
        e1.record()
        flops = 2.0 * N * Ho * Wo * kh * kw * (ctot if alg_cin is None else alg_cin) * p.Cout_store
        _PROFILE.append((e0, e1, flops, tag, nbytes, kh * kw))
        return out
    _call("fbanet_conv_gemm_sm100", p, tag=tag, nbytes=nbytes)
    return out


def profile_conv_gemm(fn, stream) -> dict:
    """Run ``fn`` once with CUDA events around every implicit-GEMM launch; returns summed device time,
    algorithmic FLOPs (2*M*K*N with unpadded channels) and the achieved TFLOP/s of that kernel family, plus the same split by
    class -- ``"conv"`` (3x3 / 4x4 taps: tensor-pipe bound, judged in TFLOP/s) and ``"gemm1x1"`` (linear layers and 1x1 / transposed
    convs: HBM bound at these channel counts, judged in GB/s) -- and per distinct launch shape (``"shapes"``: tag -> launches, ms,
    flops, algorithmic bytes)."""
    global _PROFILE
    _PROFILE = []
    try:
        fn()
        stream.synchronize()
        recs = [(a.elapsed_time(b), f, tag, nb, taps) for a, b, f, tag, nb, taps in _PROFILE]
    finally:
        _PROFILE = None
    ms = sum(r[0] for r in recs)
    flops = sum(r[1] for r in recs)
    classes, shapes = {}, {}
    for t, f, tag, nb, taps in recs:
        c = classes.setdefault("conv" if taps > 1 else "gemm1x1", {"ms": 0.0, "flops": 0.0, "bytes": 0, "launches": 0})
        c["ms"] += t; c["flops"] += f; c["bytes"] += nb; c["launches"] += 1
        sh = shapes.setdefault(tag, {"ms": 0.0, "flops": 0.0, "bytes": 0, "launches": 0, "taps": taps})
        sh["ms"] += t; sh["flops"] += f; sh["bytes"] += nb; sh["launches"] += 1
    return {"kernel": "fbanet_conv_gemm_sm100 (implicit-GEMM conv/linear, all launches of one step)", "ms": ms, "flops": flops,
            "launches": len(recs), "tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, "classes": classes, "shapes": shapes,
            "order": [taps for _, _, _, _, taps in recs]}


def tcgen05_supported(p: L.ConvParams) -> bool:
    return bool(L.load().fbanet_conv_gemm_tcgen05_supported(C.byref(p)))


def to_nhwc(x: torch.Tensor, cp: int, dtype: torch.dtype, im2col3x3: bool = False) -> torch.Tensor:
    """planar fp32 ``[frames,C,H,W]`` -> channels-last ``[frames,H,W,cp]`` (zero padded channels);
    ``im2col3x3``: channel ``(ky*3+kx)*C + c`` = 3x3 neighbourhood (zero padded), for the head conv."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    Fr, Cc, H, W = x.shape
    out = torch.empty((Fr, H, W, cp), device=x.device, dtype=dtype)
    p = L.ToNhwcParams()
    p.src, p.dst, p.dtype = x.data_ptr(), out.data_ptr(), _DT[dtype]
    p.frames, p.C, p.H, p.W, p.Cp = Fr, Cc, H, W, cp
    p.im2col3x3 = 1 if im2col3x3 else 0
    _call("fbanet_to_nhwc_sm100", p)
    return out


def head_conv(x: torch.Tensor, weight_kc: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype, M: Optional[torch.Tensor] = None,
              frames_per_burst: int = 0) -> torch.Tensor:
    """3x3 pad-1 conv from the planar fp32 burst ``[frames,C,H,W]`` to channels-last ``[frames,H,W,64]``;
    ``weight_kc`` is fp32 ``[9*C, 64]`` (k = (ky*3+kx)*C + c).

    ``M`` ``[frames,3,3]`` float64 dst->src homographies: K1 fused into K0 -- every sample the conv reads is the bilinearly warped burst
    pixel (the arithmetic of :func:`warp_burst`, bit for bit; frames ``f % frames_per_burst == 0`` are base frames and are copied), so
    the warped burst is never materialised.  Tensor-core (bf16, 64-channel) path only; elsewhere the burst is warped first."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    Fr, Cc, H, W = x.shape
    cout = weight_kc.shape[1]
    assert weight_kc.dtype == torch.float32 and weight_kc.is_contiguous() and weight_kc.shape[0] == 9 * Cc
    out = torch.empty((Fr, H, W, cout), device=x.device, dtype=dtype)
    p = L.HeadConvParams()
    p.src, p.dst, p.weight, p.bias, p.dtype = x.data_ptr(), out.data_ptr(), weight_kc.data_ptr(), bias.data_ptr(), _DT[dtype]
    p.frames, p.C, p.H, p.W, p.Cout = Fr, Cc, H, W, cout
    if M is not None:
        assert frames_per_burst > 0 and Fr % frames_per_burst == 0
        M = M.to(device=x.device, dtype=torch.float64).contiguous()
        assert M.shape == (Fr, 3, 3)
        p.M, p.frames_per_burst = M.data_ptr(), frames_per_burst
        try:
            _call("fbanet_head_conv_sm100", p, tag="+warp", nbytes=x.numel() * 4 + out.numel() * out.element_size())
        except RuntimeError as e:
            if "impl unsupported" not in str(e):
                raise
            # no fused kernel for this dtype / width: warp, then convolve
            xw = warp_burst(x.view(Fr // frames_per_burst, frames_per_burst, Cc, H, W), M.view(-1, frames_per_burst, 3, 3))
            return head_conv(xw.view(Fr, Cc, H, W), weight_kc, bias, dtype)
        return out
    _call("fbanet_head_conv_sm100", p, nbytes=x.numel() * 4 + out.numel() * out.element_size())
    return out


def assemble(sr: torch.Tensor, base: torch.Tensor, C_out: int, lo_offset: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[n,c] = sr[n,:,:,c] + bilinear_x4(base[n,c])``: channels-last SR ``[N,4h,4w,Cp]`` + planar fp32 base
    ``[N,C,h,w]`` (a view of the burst's frame 0) -> planar fp32 ``[N,C,4h,4w]``.  ``lo_offset > 0``: the SR value of channel
    ``c`` is ``sr[..., c] + sr[..., c + lo_offset]`` (the final conv run with hi/lo split weights)."""
    assert sr.is_cuda and sr.is_contiguous() and sr.dim() == 4
    N, H, W, Cp = sr.shape
    assert base.dtype == torch.float32 and base.shape == (N, C_out, H // 4, W // 4)
    assert base.stride(3) == 1 and base.stride(2) == W // 4 and base.stride(1) == (H // 4) * (W // 4)
    if out is None:
        out = torch.empty((N, C_out, H, W), device=sr.device, dtype=torch.float32)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (N, C_out, H, W)
    p = L.AssembleParams()
    p.sr, p.base, p.out = sr.data_ptr(), base.data_ptr(), out.data_ptr()
    p.base_img_stride = base.stride(0) if N > 1 else 0
    p.dtype, p.N, p.C, p.Cp, p.H, p.W = _DT[sr.dtype], N, C_out, Cp, H, W
    assert lo_offset == 0 or lo_offset + C_out <= Cp
    p.lo_offset = lo_offset
    _call("fbanet_assemble_sm100", p, nbytes=N * H * W * C_out * sr.element_size() + base.numel() * 4 + out.numel() * 4)
    return out


def convert_io(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """Narrow host I/O on the device: ``uint8 -> fp32`` (``x / 255``, ``train.py:82-83``), ``fp32 -> uint8`` (``clamp(x,0,1)`` then
    ``mul(255).byte()``: the reference's PNG path, ``test_in_any_resolution.py:93-101``) or ``fp32 -> fp16``; contiguous tensors of the
    same shape with a multiple of 16 elements."""
    assert src.is_cuda and dst.is_cuda and src.is_contiguous() and dst.is_contiguous() and src.shape == dst.shape
    mode = {(torch.uint8, torch.float32): L.CONVERT_U8_TO_F32, (torch.float32, torch.uint8): L.CONVERT_F32_TO_U8,
            (torch.float32, torch.float16): L.CONVERT_F32_TO_F16}[(src.dtype, dst.dtype)]
    p = L.ConvertIoParams()
    p.src, p.dst, p.n, p.mode = src.data_ptr(), dst.data_ptr(), src.numel(), mode
    _call("fbanet_convert_io_sm100", p, nbytes=src.numel() * src.element_size() + dst.numel() * dst.element_size())
    return dst


def space_to_depth(x: torch.Tensor) -> torch.Tensor:
    """channels-last view ``[N,H,W,C]`` -> contiguous ``[N,H/2,W/2,4C]``, channel ``(ys*2+xs)*C + c``."""
    ptr, Cc, ld, istr = _cl(x)
    N, H, W, _ = x.shape
    out = torch.empty((N, H // 2, W // 2, 4 * Cc), device=x.device, dtype=x.dtype)
    p = L.S2dParams()
    p.src, p.dst, p.img_stride, p.dtype = ptr, out.data_ptr(), istr, _DT[x.dtype]
    p.N, p.H, p.W, p.C, p.ld = N, H, W, Cc, ld
    _call("fbanet_space_to_depth_sm100", p)
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last dim of a ``[rows, C]`` view (row stride may exceed C)."""
    assert x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    rows, Cc = x.shape
    y = torch.empty((rows, Cc), device=x.device, dtype=x.dtype)
    p = L.LayerNormParams()
    p.x, p.y, p.gamma, p.beta = x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    p.rows, p.C, p.x_ld, p.y_ld, p.dtype, p.eps = rows, Cc, x.stride(0), Cc, _DT[x.dtype], eps
    _call("fbanet_layernorm_sm100", p, nbytes=2 * x.numel() * x.element_size())
    return y


def row_stats(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm statistics only: ``[rows, C]`` bf16 -> fp32 ``[rows, 2]`` = (mean, 1/sqrt(var + eps)); the normalisation itself is
    applied by the consumer GEMM (``conv_gemm(..., ln_stats=)``), so the normalised tensor never exists in HBM."""
    assert x.is_cuda and x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16
    rows, Cc = x.shape
    st = torch.empty((rows, 2), device=x.device, dtype=torch.float32)
    p = L.LayerNormParams()
    p.x, p.stats = x.data_ptr(), st.data_ptr()
    p.rows, p.C, p.x_ld, p.y_ld, p.dtype, p.eps = rows, Cc, x.stride(0), Cc, _DT[x.dtype], eps
    _call("fbanet_layernorm_sm100", p, nbytes=x.numel() * x.element_size() + st.numel() * 4)
    return st


def fold_layernorm(w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dtype: torch.dtype):
    """Fold ``LN(x) = (x - mean) * rstd * gamma + beta`` into the linear layer ``y = W LN(x) + b`` that consumes it.

    With ``W' = W diag(gamma)`` and every row of ``W'`` centred, ``W'' = W' - rowmean(W') 1^T``, one has ``W'' x = W' (x - mean(x) 1)``
    (``mean(x) = 1^T x / K``): the mean subtraction moves into the weights, and ``y = rstd * (W'' x) + (W beta + b)``.
    Returns (``W''`` in ``dtype``, fp32 bias ``W beta + b``).  Centring is done in fp64; rounding to ``dtype`` then leaves row sums of
    order sqrt(K) * 2^-10 * |W|, which would leak ``mean(x)`` into the output (activations with |mean| >> sigma: measurable), so the
    residual of every row is pushed into its smallest-magnitude elements (whose ulps are far finer)."""
    w64, g64, be64 = w.detach().double(), gamma.detach().double(), beta.detach().double()
    wg = w64 * g64[None, :]
    wf = (wg - wg.mean(1, keepdim=True)).to(dtype)
    # pass 1: the residual is shared by the 8 smallest-magnitude elements of the row (each moves by about one typical ulp);
    # pass 2: what their own rounding leaves (~2e-5) goes into the single smallest element.
    for t in (8, 1):
        r = wf.double().sum(1, keepdim=True)
        idx = wf.abs().float().topk(min(t, wf.shape[1]), dim=1, largest=False).indices
        wf.scatter_(1, idx, (wf.gather(1, idx).double() - r / idx.shape[1]).to(dtype))
    bias = (w64 @ be64 + b.detach().double()).float().contiguous()
    return wf.contiguous(), bias


def round_rowsum(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Round the rows of a GEMM weight matrix ``[rows, K]`` to ``dtype`` so that every row keeps its SUM.

    Plain round-to-nearest leaves each output channel with a row-sum error of order ``sqrt(K) * 2^-10 * |w|``, i.e. a fixed bias
    proportional to the COMMON MODE of its inputs -- and most inputs here have one (GELU / ReLU / PReLU outputs are mostly positive,
    attention outputs are averages).  That bias is the same at every pixel, so it is the part of the weight-rounding error that moves
    the PSNR against a ground truth, the north-star's bf16 criterion (``tools/emulate_bf16.py``: the PSNR delta of the parity test is
    entirely weight rounding; sum-preserving rows lower it on every seed, mean 0.0048 -> 0.0035 dB).  The residual of a row is pushed
    into its 8 smallest-magnitude non-zero elements (about one typical ulp each; their own ulps are far finer), then what their
    rounding leaves into the single smallest (the scheme :func:`fold_layernorm` uses for its zero row sums).  fp32: a plain cast."""
    if dtype != torch.bfloat16:
        return w.detach().to(dtype).contiguous()
    w64 = w.detach().double()
    wf = w64.to(dtype)
    mag = wf.abs().float()
    mag[mag == 0] = float("inf")          # zero columns (channel padding) multiply zero inputs: leave them alone
    for t in (8, 1):
        r = wf.double().sum(1, keepdim=True) - w64.sum(1, keepdim=True)
        idx = mag.topk(min(t, wf.shape[1]), dim=1, largest=False).indices
        wf.scatter_(1, idx, (wf.gather(1, idx).double() - r / idx.shape[1]).to(dtype))
    return wf.contiguous()


def expand_rel_pos_bias(bias_table: torch.Tensor, win: int) -> torch.Tensor:
    """Dense ``[heads, win^2, NP]`` fp32 bias in log2 units for the tensor-core attention kernel (``NP`` = keys padded to
    a multiple of 16, padded keys = -1e30): ``log2(e) * table[(yi-yj+w-1)(2w-1) + (xi-xj+w-1), h]``."""
    N = win * win
    NP = (N + 15) // 16 * 16
    ys, xs = torch.meshgrid(torch.arange(win), torch.arange(win), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    idx = ((ys[:, None] - ys[None, :] + win - 1) * (2 * win - 1) + (xs[:, None] - xs[None, :] + win - 1)).to(bias_table.device)
    dense = bias_table.float()[idx.view(-1)].view(N, N, -1).permute(2, 0, 1) * 1.4426950408889634
    out = torch.full((dense.shape[0], N, NP), -1e30, device=bias_table.device, dtype=torch.float32)
    out[:, :, :N] = dense
    return out.contiguous()


def expand_rel_pos_bias_wrap(bias_expanded: torch.Tensor, win: int) -> torch.Tensor:
    """Tables of the tcgen05 attention kernel for the windows of a SHIFTED layer that wrap around the image edge: ``[heads, 3, N, NP]``
    fp32 for wrap type 1 (bottom edge), 2 (right edge), 3 (corner).  The kernel fetches such a window as 2 / 4 TMA boxes, so its tile
    holds the tokens in box order -- type 1: rows ``iy < 5`` then ``iy >= 5`` (= natural order); type 2: columns ``ix < 5`` of every row,
    then ``ix >= 5``; type 3: the four 5 x 5 quadrants -- and the table is the bias with rows and columns permuted alike, plus the Swin
    shift mask (``-100`` in the reference, ``layers/fba_net.py:151-184``; log2 units here) between tokens of different boxes, which
    are exactly the tokens of different shift regions."""
    heads, N, NP = bias_expanded.shape
    half = win // 2
    dev = bias_expanded.device
    r = torch.arange(N, device=dev)
    toks, blocks = [], []
    for t in (1, 2, 3):
        if t == 1:
            tok, blk = r, r // (half * win)
        elif t == 2:
            rr = r % (half * win)
            tok, blk = (rr // half) * win + (r // (half * win)) * half + rr % half, r // (half * win)
        else:
            qd, rr = r // (half * half), r % (half * half)
            tok, blk = ((qd // 2) * half + rr // half) * win + (qd % 2) * half + rr % half, qd
        toks.append(tok), blocks.append(blk)
    out = torch.full((heads, 3, N, NP), -1e30, device=dev, dtype=torch.float32)
    for i, (tok, blk) in enumerate(zip(toks, blocks)):
        tbl = bias_expanded[:, tok][:, :, tok]
        out[:, i, :, :N] = tbl + (blk[:, None] != blk[None, :]).float() * (-100.0 * 1.4426950408889634)
    return out.contiguous()


def window_attention(qkv: torch.Tensor, bias_table: torch.Tensor, B: int, H: int, W: int, heads: int, win: int, shift: int,
                     scale: float, impl: int = L.IMPL_AUTO, bias_expanded: Optional[torch.Tensor] = None,
                     q_prescaled: bool = False, bias_wrap: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv ``[B*H*W, 3C]`` -> ``[B*H*W, C]``.  ``q_prescaled``: the q columns were produced by projection weights that
    already carry ``scale * log2(e)`` (bf16 tensor-core kernels only); ``scale`` is then ignored."""
    assert qkv.is_cuda and qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[0] == B * H * W
    Cc = qkv.shape[1] // 3
    out = torch.empty((B * H * W, Cc), device=qkv.device, dtype=qkv.dtype)
    assert bias_table.dtype == torch.float32 and bias_table.is_contiguous() and bias_table.shape == ((2 * win - 1) ** 2, heads)
    p = L.AttnParams()
    p.qkv, p.out, p.bias_table, p.dtype = qkv.data_ptr(), out.data_ptr(), bias_table.data_ptr(), _DT[qkv.dtype]
    p.B, p.H, p.W, p.C, p.heads, p.win, p.shift = B, H, W, Cc, heads, win, shift
    p.qkv_ld, p.out_ld, p.scale, p.impl = qkv.stride(0), Cc, scale, impl
    p.q_prescaled = 1 if q_prescaled else 0
    if bias_expanded is not None:
        N = win * win
        assert bias_expanded.dtype == torch.float32 and bias_expanded.is_contiguous() and bias_expanded.shape == (heads, N, (N + 15) // 16 * 16)
        p.bias_expanded = bias_expanded.data_ptr()
    if bias_wrap is not None:
        assert bias_expanded is not None and bias_wrap.dtype == torch.float32 and bias_wrap.is_contiguous() and bias_wrap.shape == (heads, 3) + tuple(bias_expanded.shape[1:])
        p.bias_wrap = bias_wrap.data_ptr()
    global LAST_ATTENTION_ON_TCGEN05
    LAST_ATTENTION_ON_TCGEN05 = bool(L.load().fbanet_window_attention_tcgen05_supported(C.byref(p))) and impl != L.IMPL_SIMT
    _call("fbanet_window_attention_sm100", p, tag=f"dh{Cc // heads} {'tcgen05' if LAST_ATTENTION_ON_TCGEN05 else 'mma.sync/simt'}",
          nbytes=(qkv.numel() + out.numel()) * qkv.element_size())
    return out


def dwconv3x3(x: torch.Tensor, weight9c: torch.Tensor, bias: torch.Tensor, act: int) -> torch.Tensor:
    """depthwise 3x3 pad 1 + bias + act on contiguous ``[N,H,W,C]``; ``weight9c`` is ``[9,C]`` fp32."""
    assert x.is_cuda and x.is_contiguous() and x.dim() == 4
    N, H, W, Cc = x.shape
    y = torch.empty_like(x)
    p = L.DwconvParams()
    p.x, p.y, p.weight, p.bias, p.dtype = x.data_ptr(), y.data_ptr(), weight9c.data_ptr(), bias.data_ptr(), _DT[x.dtype]
    p.N, p.H, p.W, p.C, p.act = N, H, W, Cc, act
    _call("fbanet_dwconv3x3_sm100", p)
    return y


def leff_fc2(h1: torch.Tensor, dw_w9c: torch.Tensor, dw_b: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, out: torch.Tensor,
             residual: Optional[torch.Tensor], act: int) -> Optional[torch.Tensor]:
    """Fused LeFF tail (bf16): ``out = Linear2(act(depthwise3x3(h1) + dw_b)) + b2 + residual``.  ``h1`` contiguous
    ``[N,H,W,Hd]``; ``w2`` ``[C,Hd]`` bf16 -- or both fp16 (``conv_gemm`` with an fp16 ``out``): half2 depthwise arithmetic.  Returns ``None`` when the fused kernel does not take the shape (caller
    then runs dwconv + GEMM)."""
    assert h1.is_cuda and h1.is_contiguous() and h1.dtype in (torch.bfloat16, torch.float16) and h1.dim() == 4
    N, H, W, Hd = h1.shape
    Cc = w2.shape[0]
    assert w2.dtype == h1.dtype and w2.is_contiguous() and w2.shape == (Cc, Hd) and out.shape == (N, H, W, Cc)
    p = L.LeffFc2Params()
    p.f16 = 1 if h1.dtype == torch.float16 else 0   # fp16 hidden map + fp16 fc2 weights: half2 depthwise producer (tanh GELU)
    p.h1, p.dw_weight, p.dw_bias, p.w2, p.bias2 = h1.data_ptr(), dw_w9c.data_ptr(), dw_b.data_ptr(), w2.data_ptr(), b2.data_ptr()
    op, _, old, ois = _cl(out)
    p.out, p.out_ld, p.out_img_stride = op, old, ois
    if residual is not None:
        rp, _, rld, ris = _cl(residual)
        p.residual, p.res_ld, p.res_img_stride = rp, rld, ris
    p.N, p.H, p.W, p.C, p.Hd, p.act = N, H, W, Cc, Hd, act
    if not L.load().fbanet_leff_fc2_supported(C.byref(p)):
        return None
    _call("fbanet_leff_fc2_sm100", p, tag=f"{Hd}->{Cc} @{H}x{W}", nbytes=(h1.numel() + out.numel() * (2 if residual is not None else 1) + w2.numel()) * 2)
    return out


def leff_mlp(x: torch.Tensor, w1h: torch.Tensor, b1h: torch.Tensor, dw_w9c_h: torch.Tensor, dw_bh: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
             out: torch.Tensor, residual: Optional[torch.Tensor], act: int) -> Optional[torch.Tensor]:
    """The whole LeFF MLP in one kernel (bf16): ``out = Linear2(act(depthwise3x3(act(Linear1(x))) + dw_b)) + b2 + residual``
    (``layers/locally_enhanced_feed_forward.py:25-57``); the 4C-channel hidden map never goes to HBM.  ``x`` / ``out`` /
    ``residual``: channels-last views ``[N,H,W,C]``.  ``w1h [Hd,C]`` bf16, ``b1h [Hd]``, ``dw_w9c_h [9,Hd]``, ``dw_bh [Hd]`` hold HALF the
    layer's values (the kernel's contract, see ``include/fbanet_b200.h``); ``w2 [C,Hd]`` bf16 -- or fp16, which makes the kernel keep its
    on-chip hidden tile in fp16 and run the depthwise conv + GELU on packed half2 (tanh GELU only) --, ``b2 [C]``.  Returns ``None`` when
    the kernel does not take the shape (C > 128)."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
    N, H, W, Cc = x.shape
    Hd = w1h.shape[0]
    assert w1h.dtype == torch.bfloat16 and w1h.is_contiguous() and w1h.shape == (Hd, Cc)
    assert w2.dtype in (torch.bfloat16, torch.float16) and w2.is_contiguous() and w2.shape == (Cc, Hd) and out.shape == (N, H, W, Cc)
    assert dw_w9c_h.shape == (9, Hd) and dw_w9c_h.dtype == torch.float32 and dw_w9c_h.is_contiguous()
    p = L.LeffMlpParams()
    p.w2_f16 = 1 if w2.dtype == torch.float16 else 0   # fp16 fc2 weights select the fp16 hidden tile / half2 depthwise path (tanh GELU)
    p.w1, p.bias1, p.dw_weight, p.dw_bias, p.w2, p.bias2 = (w1h.data_ptr(), b1h.data_ptr(), dw_w9c_h.data_ptr(), dw_bh.data_ptr(), w2.data_ptr(),
                                                          b2.data_ptr())
    xp, _, xld, xis = _cl(x)
    p.x, p.x_ld, p.x_img_stride = xp, xld, xis
    op, _, old, ois = _cl(out)
    p.out, p.out_ld, p.out_img_stride = op, old, ois
    if residual is not None:
        rp, _, rld, ris = _cl(residual)
        p.residual, p.res_ld, p.res_img_stride = rp, rld, ris
    p.N, p.H, p.W, p.C, p.Hd, p.act = N, H, W, Cc, Hd, act
    if not L.load().fbanet_leff_mlp_supported(C.byref(p)):
        return None
    _call("fbanet_leff_mlp_sm100", p, tag=f"{Cc}->{Hd}->{Cc} @{H}x{W}",
          nbytes=(x.numel() + out.numel() * (2 if residual is not None else 1) + w1h.numel() + w2.numel()) * 2)
    return out


def faf_gate(feat: torch.Tensor, wsum: torch.Tensor, want_gate: bool = True, want_gated: bool = False,
             score: Optional[torch.Tensor] = None):
    """feat ``[B,F,H,W,C]`` contiguous -> gate ``[B,F-1,H,W]`` fp32 and/or gated features
    ``[B,H,W,F*C]`` (pixel-major; frame 0 copied, frames >= 1 scaled) for the tensor-core fusion GEMM.
    ``score`` ``[B*F,H,W,2]`` fp32: the ``wsum`` dot products already computed by :func:`faf_scores`."""
    assert feat.is_cuda and feat.is_contiguous() and feat.dim() == 5
    B, Fr, H, W, Cc = feat.shape
    gate = torch.empty((B, Fr - 1, H, W), device=feat.device, dtype=torch.float32) if want_gate else None
    gated = torch.empty((B, H, W, Fr * Cc), device=feat.device, dtype=feat.dtype) if want_gated else None
    p = L.FafGateParams()
    p.feat, p.wsum, p.dtype = feat.data_ptr(), wsum.data_ptr(), _DT[feat.dtype]
    p.gate = gate.data_ptr() if gate is not None else None
    p.gated = gated.data_ptr() if gated is not None else None
    p.B, p.F, p.H, p.W, p.C = B, Fr, H, W, Cc
    if score is not None:
        assert score.dtype == torch.float32 and score.is_contiguous() and score.shape == (B * Fr, H, W, 2)
        p.score = score.data_ptr()
    nbytes = feat.numel() * feat.element_size() * (2 if want_gated else 1) + (gate.numel() * 4 if gate is not None else 0) + (score.numel() * 4 if score is not None else 0)
    _call("fbanet_faf_gate_sm100", p, nbytes=nbytes)
    if want_gate and want_gated:
        return gate, gated
    return gated if want_gated else gate


def faf_fuse_score_weight(wsum: torch.Tensor) -> torch.Tensor:
    """Gate weights ``wsum [9,64]`` fp32 as the tap-stacked B operand of :func:`faf_fuse`: bf16 ``[32,64]``, rows ``2t`` / ``2t+1`` = hi / lo
    bf16 halves of tap ``t`` (their sum carries ~16 mantissa bits of the fp32 weights), rows 18..31 zero."""
    hi = wsum.to(torch.bfloat16)
    lo = (wsum - hi.float()).to(torch.bfloat16)
    w = torch.zeros((32, wsum.shape[1]), device=wsum.device, dtype=torch.bfloat16)
    w[0:18:2], w[1:18:2] = hi, lo
    return w.contiguous()


def faf_fuse(feat: torch.Tensor, score_weight: torch.Tensor, fuse_weight: torch.Tensor, bias: torch.Tensor, alpha: torch.Tensor,
             out: torch.Tensor, want_gate: bool = False):
    """K2 in one pass (bf16): FAF gate + the ``F*64 -> 64`` 1x1 fusion conv + PReLU (``blocks/federated_affinity_fusion.py:79-105,
    121-128``) from ``feat [B,F,H,W,64]``, every feature read from HBM once; ``out``: channels-last view ``[B,H,W,64]``.  Returns
    ``(out, gate or None)``, or ``None`` when the kernel does not take the shape."""
    assert feat.is_cuda and feat.is_contiguous() and feat.dtype == torch.bfloat16 and feat.dim() == 5
    B, Fr, H, W, Cc = feat.shape
    assert score_weight.shape == (32, Cc) and score_weight.dtype == torch.bfloat16 and score_weight.is_contiguous()
    assert fuse_weight.shape == (Cc, Fr * Cc) and fuse_weight.dtype == torch.bfloat16 and fuse_weight.is_contiguous()
    assert out.shape == (B, H, W, Cc) and out.dtype == torch.bfloat16
    gate = torch.empty((B, Fr - 1, H, W), device=feat.device, dtype=torch.float32) if want_gate else None
    p = L.FafFuseParams()
    p.feat, p.score_weight, p.fuse_weight, p.bias, p.alpha = feat.data_ptr(), score_weight.data_ptr(), fuse_weight.data_ptr(), bias.data_ptr(), alpha.data_ptr()
    p.gate = gate.data_ptr() if gate is not None else None
    op, _, old, ois = _cl(out)
    p.out, p.out_ld, p.out_img_stride = op, old, ois
    p.B, p.F, p.H, p.W, p.C = B, Fr, H, W, Cc
    if not L.load().fbanet_faf_fuse_supported(C.byref(p)):
        return None
    # algorithmic bytes by SURVEY 8(d): (2F+2) H W E s -- what the three-launch form moves (read feat, write + read gated, write z); this
    # kernel itself moves (F+1) H W E s
    _call("fbanet_faf_fuse_sm100", p, nbytes=(Fr + 1) * B * H * W * Cc * 2 + (gate.numel() * 4 if gate is not None else 0))
    return out, gate


def faf_score_weight(wsum: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Pack the summed FAF kernel ``wsum [9][C]`` (fp32) as the ``[16, 9*C]`` weight of a 3x3 implicit GEMM whose output
    column 0 / 1 carry the hi / lo ``dtype`` halves of ``wsum`` (hi + lo keeps ~16 mantissa bits), rows 2..15 zero."""
    w = wsum.reshape(-1).float()
    hi = w.to(dtype)
    lo = (w - hi.float()).to(dtype)
    out = torch.zeros((16, w.numel()), device=wsum.device, dtype=dtype)
    out[0], out[1] = hi, lo
    return out.contiguous()


def faf_scores(feat: torch.Tensor, score_weight: torch.Tensor) -> torch.Tensor:
    """``wsum (*) feat`` for every frame on the tensor cores: feat ``[B,F,H,W,C]`` bf16 -> ``[B*F,H,W,2]`` fp32 (hi, lo parts)."""
    B, Fr, H, W, Cc = feat.shape
    out = torch.empty((B * Fr, H, W, 2), device=feat.device, dtype=torch.float32)
    return conv_gemm([feat.view(B * Fr, H, W, Cc)], score_weight, out, kh=3, kw=3, pad=1, store_mode=L.STORE_NHWC_F32, cout_store=2,
                     impl=L.IMPL_TCGEN05)


def warp_burst(burst: torch.Tensor, M: torch.Tensor, layout: str = "BTCHW", return_coords: bool = False):
    """Homography warp with bilinear sampling (cv2 ``INTER_LINEAR + WARP_INVERSE_MAP``, border 0).

    ``burst``: fp32 ``[B,T,C,H,W]`` (layout ``"BTCHW"``) or ``[B,T,H,W,C]`` (``"BTHWC"``, the cv2 layout);
    ``M``: ``[B,T,3,3]`` float64 dst->src matrices (entry for frame 0 ignored: the base frame is copied).
    """
    assert burst.is_cuda and burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 5
    B, T = burst.shape[:2]
    if layout == "BTCHW":
        Cc, H, W = burst.shape[2:]
        sf, sc, sy, sx = Cc * H * W, H * W, W, 1
    elif layout == "BTHWC":
        H, W, Cc = burst.shape[2:]
        sf, sy, sx, sc = H * W * Cc, W * Cc, Cc, 1
    else:
        raise ValueError(layout)
    M = M.to(device=burst.device, dtype=torch.float64).contiguous()
    assert M.shape == (B, T, 3, 3)
    out = torch.empty_like(burst)
    coords = torch.empty((B, T, H, W, 2), device=burst.device, dtype=torch.float64) if return_coords else None
    p = L.WarpParams()
    p.src, p.dst, p.M = burst.data_ptr(), out.data_ptr(), M.data_ptr()
    p.coords = coords.data_ptr() if coords is not None else None
    p.s_frame, p.s_y, p.s_x, p.s_c = sf, sy, sx, sc
    p.d_frame, p.d_y, p.d_x, p.d_c = sf, sy, sx, sc
    p.frames, p.frames_per_burst, p.H, p.W, p.C = B * T, T, H, W, Cc
    _call("fbanet_warp_sm100", p, nbytes=2 * burst.numel() * 4)
    return (out, coords) if return_coords else out


def flow_warp_burst(burst: torch.Tensor, flow: torch.Tensor, layout: str = "BTCHW") -> torch.Tensor:
    """Optical-flow registration (``registration/optical_flow/register.py:11-47``): every non-base frame is resampled at
    ``grid - flow`` bilinearly with edge-clamped indices (``map_coordinates(order=1, mode="nearest")``).

    ``burst``: fp32 ``[B,T,C,H,W]`` (``"BTCHW"``) or ``[B,T,H,W,C]`` (``"BTHWC"``, the reference's frame layout);
    ``flow``: fp32 ``[B,T-1,H,W,2]`` with last axis (dy, dx) -- no field for the base frame, which is copied.
    ``T = 1`` bursts with ``flow [B,1,H,W,2]`` are the single-frame ``register_frame(frame, flow)`` call."""
    assert burst.is_cuda and burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 5
    B, T = burst.shape[:2]
    if layout == "BTCHW":
        Cc, H, W = burst.shape[2:]
        sf, sc, sy, sx = Cc * H * W, H * W, W, 1
    elif layout == "BTHWC":
        H, W, Cc = burst.shape[2:]
        sf, sy, sx, sc = H * W * Cc, W * Cc, Cc, 1
    else:
        raise ValueError(layout)
    single = T == 1
    flow = flow.to(device=burst.device, dtype=torch.float32).contiguous()
    assert flow.shape == (B, 1 if single else T - 1, H, W, 2), f"flow {tuple(flow.shape)} does not match burst {tuple(burst.shape)}"
    out = torch.empty_like(burst)
    p = L.FlowWarpParams()
    p.src, p.dst, p.flow = burst.data_ptr(), out.data_ptr(), flow.data_ptr()
    p.s_frame, p.s_y, p.s_x, p.s_c = sf, sy, sx, sc
    p.d_frame, p.d_y, p.d_x, p.d_c = sf, sy, sx, sc
    p.frames, p.frames_per_burst, p.H, p.W, p.C = B * T, 0 if single else T, H, W, Cc
    _call("fbanet_flow_warp_sm100", p, nbytes=2 * burst.numel() * 4 + flow.numel() * 4)
    return out


def ecc_homography_burst(burst: torch.Tensor, layout: str = "BTCHW", gray_weights=None, max_iters: int = 100, eps: float = 1e-10,
                         init: Optional[torch.Tensor] = None):
    """ECC alignment of every non-base frame to frame 0 of its burst: the GPU form of ``register_frame``'s
    ``cv2.findTransformECC(gray(img1), gray(img2), eye(3), MOTION_HOMOGRAPHY, (COUNT|EPS, 100, 1e-10))``
    (``homography_alignment.py:19-45``).  Returns ``(M [B,T,3,3] float64, rho [B,T] float64, iters [B,T] int32)``; ``M`` maps
    base-frame pixel coordinates to frame coordinates -- pass it straight to :func:`warp_burst` (``WARP_INVERSE_MAP``, ``:46-55``).
    ``M[:, 0]`` is the identity.  ``iters < 0`` marks pairs where cv2 would have thrown ("stopped before its convergence").

    ``gray_weights``: per-channel weights of the grey conversion; default = ``cv2.COLOR_BGR2GRAY`` applied to the channels in
    storage order (0.114, 0.587, 0.299) for 3 channels, the channel mean otherwise."""
    assert burst.is_cuda and burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 5
    B, T = burst.shape[:2]
    assert T >= 2, "a burst needs a base frame and at least one frame to align"
    if layout == "BTCHW":
        Cc, H, W = burst.shape[2:]
        sf, sc, sy, sx = Cc * H * W, H * W, W, 1
    elif layout == "BTHWC":
        H, W, Cc = burst.shape[2:]
        sf, sy, sx, sc = H * W * Cc, W * Cc, Cc, 1
    else:
        raise ValueError(layout)
    if gray_weights is None:
        gray_weights = (0.114, 0.587, 0.299) if Cc == 3 else tuple(1.0 / Cc for _ in range(Cc))
    assert len(gray_weights) == Cc <= 4
    dev = burst.device
    planes = torch.empty((B * T, 3, H, W), device=dev, dtype=torch.float32)
    pp = L.EccPrepareParams()
    pp.src, pp.planes = burst.data_ptr(), planes.data_ptr()
    pp.s_frame, pp.s_y, pp.s_x, pp.s_c = sf, sy, sx, sc
    for c in range(4):
        pp.gray_weight[c] = float(gray_weights[c]) if c < Cc else 0.0
    pp.frames, pp.H, pp.W, pp.C = B * T, H, W, Cc
    _call("fbanet_ecc_prepare_sm100", pp, nbytes=(burst.numel() + planes.numel()) * 4)
    if init is None:
        M = torch.eye(3, dtype=torch.float64, device=dev).repeat(B, T, 1, 1).contiguous()
    else:
        M = init.to(device=dev, dtype=torch.float64).contiguous().clone()
        assert M.shape == (B, T, 3, 3)
    rho = torch.zeros((B, T), dtype=torch.float64, device=dev)
    rho[:, 0] = 1.0
    iters = torch.zeros((B, T), dtype=torch.int32, device=dev)
    p = L.EccParams()
    p.planes, p.warp, p.rho, p.iters_done = planes.data_ptr(), M.data_ptr(), rho.data_ptr(), iters.data_ptr()
    p.eps, p.frames, p.frames_per_burst, p.H, p.W, p.max_iters = float(eps), B * T, T, H, W, int(max_iters)
    _call("fbanet_ecc_homography_sm100", p)
    return M, rho, iters


def training_loss(restored: torch.Tensor, target: torch.Tensor, eps: float = 1e-3, gw_weight: float = 3.0, need_grad: bool = True,
                  clamp_restored: bool = False):
    """Training loss of the reference trainer (``train.py.bak:118-119,168``): ``CharbonnierLoss()(restored, target) + 3 *
    GWLoss()(restored, target)`` (``losses.py:39-80``) and its gradient with respect to ``restored`` in one kernel pass.
    ``clamp_restored``: the trainer's ``restored = torch.clamp(restored, 0, 1)`` (``train.py.bak:167``) in front of BOTH criteria
    (``train_step`` turns it on); off = the bare ``losses.py`` criteria.

    ``restored``, ``target``: fp32 ``[B,C,H,W]`` on the GPU.  Returns ``(loss [3] float64 = (total, charbonnier, gw), grad or None)``."""
    assert restored.is_cuda and restored.dtype == torch.float32 and restored.is_contiguous() and restored.dim() == 4
    assert target.shape == restored.shape and target.dtype == torch.float32 and target.is_contiguous() and target.device == restored.device
    B, Cc, H, W = restored.shape
    lib = L.load()
    nws = lib.fbanet_train_loss_workspace_doubles(B * Cc, H, W)
    ws = torch.empty(nws, dtype=torch.float64, device=restored.device)
    loss = torch.empty(3, dtype=torch.float64, device=restored.device)
    grad = torch.empty_like(restored) if need_grad else None
    p = L.TrainLossParams()
    p.x, p.y, p.partial, p.loss = restored.data_ptr(), target.data_ptr(), ws.data_ptr(), loss.data_ptr()
    p.grad = grad.data_ptr() if grad is not None else None
    p.eps, p.gw_weight, p.inv_n = float(eps), float(gw_weight), 1.0 / restored.numel()
    p.planes, p.H, p.W, p.clamp_restored = B * Cc, H, W, int(bool(clamp_restored))
    _call("fbanet_train_loss_sm100", p, nbytes=restored.numel() * 4 * (3 if need_grad else 2))
    return loss, grad


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, lr: float,
              betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, decoupled: bool = True, grad_scale: float = 1.0) -> None:
    """One ``torch.optim.AdamW`` (``decoupled``) / ``Adam`` update (``train.py.bak:72-78``) over flat fp32 buffers, in place.
    ``step`` is the 1-based step count after this update (torch's ``state["step"]`` once incremented)."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    assert step >= 1
    p = L.AdamParams()
    p.param, p.grad, p.exp_avg, p.exp_avg_sq, p.n = param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel()
    p.lr, p.beta1, p.beta2, p.eps, p.weight_decay = float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay)
    p.step_size = float(lr) / (1.0 - float(betas[0]) ** step)
    p.bias2_sqrt = math.sqrt(1.0 - float(betas[1]) ** step)
    p.one_minus_beta1, p.one_minus_beta2 = 1.0 - float(betas[0]), 1.0 - float(betas[1])
    p.grad_scale, p.decoupled = float(grad_scale), 1 if decoupled else 0
    _call("fbanet_adam_step_sm100", p, nbytes=param.numel() * 28)


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, kh: int = 1, kw: int = 1, stride: int = 1, pad: int = 0, need_bias: bool = True,
               dw: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None, accumulate: bool = False, splits: Optional[int] = None):
    """Weight and bias gradient of ``y = conv(x, w) + b`` (or of a linear layer: ``kh = kw = 1``): ``x`` ``[N,H,W,Cin]``, ``dy``
    ``[N,Ho,Wo,Cout]`` channels-last views of one dtype -> ``(dw [Cout,Cin,kh,kw] fp32, db [Cout] fp32 or None)`` in the torch
    parameter layouts (so they can be views into ``train.FlatParams.grad``).  Fixed-order split reduction: bit-reproducible."""
    assert x.is_cuda and dy.is_cuda and x.dtype == dy.dtype
    N, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    assert dy.shape[:3] == (N, Ho, Wo), (dy.shape, (N, Ho, Wo))
    Cout = dy.shape[3]
    xp, _, xld, xis = _cl(x)
    dp, _, dld, dis = _cl(dy)
    if dw is None:
        dw = torch.empty((Cout, Cin, kh, kw), device=x.device, dtype=torch.float32)
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == Cout * Cin * kh * kw
    if need_bias and db is None:
        db = torch.empty(Cout, device=x.device, dtype=torch.float32)
    if db is not None:
        assert db.dtype == torch.float32 and db.is_contiguous() and db.numel() == Cout
    P = N * Ho * Wo
    if splits is None:   # enough CTAs for two waves of the 148 SMs, at least 256 pixels per chunk
        tiles = ((Cout + 63) // 64) * kh * kw * ((Cin + 63) // 64)
        splits = max(1, min((2 * 148 + tiles - 1) // tiles, (P + 255) // 256, 1024))
    partial = torch.empty(splits * (Cout * kh * kw * Cin + Cout), device=x.device, dtype=torch.float32)
    p = L.WgradParams()
    p.x, p.dy, p.dw, p.partial = xp, dp, dw.data_ptr(), partial.data_ptr()
    p.db = db.data_ptr() if db is not None else None
    p.x_img_stride, p.dy_img_stride, p.x_ld, p.dy_ld = xis, dis, xld, dld
    p.dtype, p.N, p.H, p.W, p.Cin, p.Ho, p.Wo, p.Cout = _DT[x.dtype], N, H, W, Cin, Ho, Wo, Cout
    p.KH, p.KW, p.stride, p.pad, p.splits, p.accumulate = kh, kw, stride, pad, splits, 1 if accumulate else 0
    esz = x.element_size()
    _call("fbanet_wgrad_sm100", p, tag=f"wgrad k{kh}s{stride} {Cin}->{Cout} @{Ho}x{Wo}", nbytes=(kh * kw * x.numel() + dy.numel()) * esz)
    return dw, db


def layernorm_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5, dgamma: Optional[torch.Tensor] = None,
                       dbeta: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Backward of the per-token LayerNorm (``layers/fba_net.py:196,246``): ``x``, ``dy`` ``[rows, C]`` contiguous, ``C <= 256`` ->
    ``(dx [rows, C] in x's dtype, dgamma [C] fp32, dbeta [C] fp32)``."""
    assert x.is_cuda and x.dim() == 2 and x.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype and dy.is_contiguous()
    rows, Cc = x.shape
    assert gamma.dtype == torch.float32 and gamma.is_contiguous() and gamma.numel() == Cc and Cc <= 256
    dx = torch.empty_like(x)
    dgamma = torch.empty(Cc, device=x.device, dtype=torch.float32) if dgamma is None else dgamma
    dbeta = torch.empty(Cc, device=x.device, dtype=torch.float32) if dbeta is None else dbeta
    for t in (dgamma, dbeta):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == Cc
    blocks = L.load().fbanet_layernorm_bwd_blocks(rows)
    partial = torch.empty(blocks * 2 * Cc, device=x.device, dtype=torch.float32)
    p = L.LayerNormBwdParams()
    p.x, p.dy, p.gamma, p.dx = x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), dx.data_ptr()
    p.dgamma, p.dbeta, p.partial = dgamma.data_ptr(), dbeta.data_ptr(), partial.data_ptr()
    p.rows, p.eps, p.dtype, p.C, p.accumulate = rows, float(eps), _DT[x.dtype], Cc, 1 if accumulate else 0
    _call("fbanet_layernorm_bwd_sm100", p, nbytes=3 * x.numel() * x.element_size())
    return dx, dgamma, dbeta


def act_backward(x: torch.Tensor, dy: torch.Tensor, act: int, alpha: Optional[torch.Tensor] = None, dalpha: Optional[torch.Tensor] = None,
                 accumulate: bool = False):
    """``dx = dy * act'(x)`` for the PRE-activation ``x`` (``act``: ``L.ACT_RELU / ACT_PRELU / ACT_GELU_TANH / ACT_GELU_ERF``); for PReLU
    also the gradient of the scalar slope.  Returns ``dx`` (``(dx, dalpha)`` for PReLU)."""
    assert x.is_cuda and x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype
    dx = torch.empty_like(x)
    p = L.ActBwdParams()
    p.x, p.dy, p.dx, p.n, p.dtype, p.act, p.accumulate = x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _DT[x.dtype], act, 1 if accumulate else 0
    if act == L.ACT_PRELU:
        assert alpha is not None and alpha.dtype == torch.float32 and alpha.numel() == 1
        dalpha = torch.empty(1, device=x.device, dtype=torch.float32) if dalpha is None else dalpha
        assert dalpha.dtype == torch.float32 and dalpha.numel() == 1
        partial = torch.empty(L.load().fbanet_act_bwd_blocks(x.numel()), device=x.device, dtype=torch.float32)
        p.alpha, p.dalpha, p.partial = alpha.data_ptr(), dalpha.data_ptr(), partial.data_ptr()
    _call("fbanet_act_bwd_sm100", p, nbytes=3 * x.numel() * x.element_size())
    return (dx, dalpha) if act == L.ACT_PRELU else dx


def act_forward(x: torch.Tensor, act: int, alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Training-mode stand-alone activation ``y = act(x)``; the caller keeps ``x`` (the pre-activation) for :func:`act_backward`."""
    assert x.is_cuda and x.is_contiguous()
    y = torch.empty_like(x)
    p = L.ActFwdParams()
    p.x, p.y, p.n, p.dtype, p.act = x.data_ptr(), y.data_ptr(), x.numel(), _DT[x.dtype], act
    if act == L.ACT_PRELU:
        assert alpha is not None and alpha.dtype == torch.float32 and alpha.numel() == 1
        p.alpha = alpha.data_ptr()
    _call("fbanet_act_fwd_sm100", p, nbytes=2 * x.numel() * x.element_size())
    return y


def dwconv3x3_backward(x: torch.Tensor, dy: torch.Tensor, weight9c: torch.Tensor, need_dx: bool = True, dw: Optional[torch.Tensor] = None,
                       db: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Backward of the LeFF depthwise 3x3 (``layers/locally_enhanced_feed_forward.py:39-52``): ``x``, ``dy`` channels-last
    ``[N,H,W,C]``, ``weight9c`` the forward's packed fp32 ``[9,C]`` -> ``(dx [N,H,W,C] or None, dw [C,1,3,3] fp32, db [C] fp32)``
    (torch parameter layouts, so ``dw`` / ``db`` may be views into ``FlatParams.grad``)."""
    assert x.is_cuda and x.dim() == 4 and x.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype and dy.is_contiguous()
    N, H, W, Cc = x.shape
    assert weight9c.dtype == torch.float32 and weight9c.is_contiguous() and weight9c.shape == (9, Cc)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty((Cc, 1, 3, 3), device=x.device, dtype=torch.float32) if dw is None else dw
    db = torch.empty(Cc, device=x.device, dtype=torch.float32) if db is None else db
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == 9 * Cc and db.dtype == torch.float32 and db.is_contiguous() and db.numel() == Cc
    blocks = L.load().fbanet_dwconv_bwd_blocks(N * H * W)
    partial = torch.empty(blocks * 10 * Cc, device=x.device, dtype=torch.float32)
    p = L.DwconvBwdParams()
    p.x, p.dy, p.weight, p.dw, p.db, p.partial = x.data_ptr(), dy.data_ptr(), weight9c.data_ptr(), dw.data_ptr(), db.data_ptr(), partial.data_ptr()
    p.dx = dx.data_ptr() if dx is not None else None
    p.dtype, p.N, p.H, p.W, p.C, p.accumulate = _DT[x.dtype], N, H, W, Cc, 1 if accumulate else 0
    _call("fbanet_dwconv3x3_bwd_sm100", p, nbytes=(3 if need_dx else 2) * x.numel() * x.element_size())
    return dx, dw, db


def window_attention_backward(qkv: torch.Tensor, dout: torch.Tensor, bias_table: torch.Tensor, B: int, H: int, W: int, heads: int, win: int,
                              shift: int, scale: float, need_dbias: bool = True, dbias: Optional[torch.Tensor] = None,
                              accumulate: bool = False):
    """Backward of :func:`window_attention` (``layers/window_attention.py:159-248``): ``qkv [B*H*W, 3C]`` (UNSCALED q, as the fp32
    forward takes it), ``dout [B*H*W, C]`` -> ``(dqkv [B*H*W, 3C], dbias_table [(2win-1)^2, heads] fp32 or None)``."""
    assert qkv.is_cuda and qkv.dim() == 2 and qkv.is_contiguous() and dout.is_cuda and dout.dim() == 2 and dout.is_contiguous()
    T, C3 = qkv.shape
    Cc = C3 // 3
    assert T == B * H * W and C3 == 3 * Cc and dout.shape == (T, Cc) and dout.dtype == qkv.dtype
    R = (2 * win - 1) ** 2
    assert bias_table.dtype == torch.float32 and bias_table.is_contiguous() and bias_table.shape == (R, heads)
    dqkv = torch.empty_like(qkv)
    p = L.AttnBwdParams()
    p.qkv, p.dout, p.dqkv, p.bias_table = qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), bias_table.data_ptr()
    if need_dbias:
        dbias = torch.empty((R, heads), device=qkv.device, dtype=torch.float32) if dbias is None else dbias
        assert dbias.dtype == torch.float32 and dbias.is_contiguous() and dbias.numel() == R * heads
        n = L.load().fbanet_attn_bwd_partial_floats(B, H, W, heads, win)
        assert n > 0, "H, W must be multiples of win"
        partial = torch.empty(n, device=qkv.device, dtype=torch.float32)
        p.dbias, p.partial = dbias.data_ptr(), partial.data_ptr()
    else:
        dbias = None
    p.dtype, p.B, p.H, p.W, p.C, p.heads, p.win, p.shift = _DT[qkv.dtype], B, H, W, Cc, heads, win, shift
    p.qkv_ld, p.dout_ld, p.dqkv_ld, p.scale, p.accumulate = C3, Cc, C3, float(scale), 1 if accumulate else 0
    _call("fbanet_window_attention_bwd_sm100", p, tag=f"attn bwd dh{Cc // heads}", nbytes=(2 * qkv.numel() + dout.numel()) * qkv.element_size())
    return dqkv, dbias


def faf_gate_backward(feat: torch.Tensor, dgated: torch.Tensor, gate: torch.Tensor, score: torch.Tensor, wsum: torch.Tensor,
                      dwsum: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Backward of :func:`faf_gate` (``blocks/federated_affinity_fusion.py:79-105``): ``feat [B,F,H,W,C]``, ``dgated [B,H,W,F*C]``
    (gradient of the gated, pixel-major features), ``gate [B,F-1,H,W]`` fp32 as the forward returned it, ``score`` fp32
    ``[B*F,H,W]`` or the ``[B*F,H,W,2]`` hi / lo form of :func:`faf_scores` -> ``(dfeat [B,F,H,W,C], dwsum [9,C] fp32)``.
    ``dwsum`` is the gradient of EVERY output channel of ``temporal_attn1.weight`` (``dW1[co, ci, ky, kx] = dwsum[ky*3+kx, ci]``);
    ``temporal_attn0`` and both biases cancel out of the gate as the reference writes it and get zero (:func:`faf_weight_grads`)."""
    assert feat.is_cuda and feat.dim() == 5 and feat.is_contiguous()
    B, Fr, H, W, Cc = feat.shape
    assert dgated.is_contiguous() and dgated.dtype == feat.dtype and dgated.numel() == feat.numel() and dgated.shape[:3] == (B, H, W)
    assert gate.dtype == torch.float32 and gate.is_contiguous() and gate.shape == (B, Fr - 1, H, W)
    assert wsum.dtype == torch.float32 and wsum.is_contiguous() and wsum.shape == (9, Cc)
    assert score.dtype == torch.float32
    if score.dim() == 4 and score.shape[-1] == 2:
        score = score[..., 0] + score[..., 1]
    score = score.reshape(B, Fr, H, W).contiguous()
    dfeat = torch.empty_like(feat)
    dscore = torch.empty((B, Fr, H, W), device=feat.device, dtype=torch.float32)
    dwsum = torch.empty((9, Cc), device=feat.device, dtype=torch.float32) if dwsum is None else dwsum
    assert dwsum.dtype == torch.float32 and dwsum.is_contiguous() and dwsum.numel() == 9 * Cc
    blocks = L.load().fbanet_faf_gate_bwd_blocks(B * Fr * H * W)
    partial = torch.empty(blocks * 9 * Cc, device=feat.device, dtype=torch.float32)
    p = L.FafGateBwdParams()
    p.feat, p.dgated, p.gate, p.score, p.wsum = feat.data_ptr(), dgated.data_ptr(), gate.data_ptr(), score.data_ptr(), wsum.data_ptr()
    p.dfeat, p.dscore, p.dwsum, p.partial = dfeat.data_ptr(), dscore.data_ptr(), dwsum.data_ptr(), partial.data_ptr()
    p.dtype, p.B, p.F, p.H, p.W, p.C, p.accumulate = _DT[feat.dtype], B, Fr, H, W, Cc, 1 if accumulate else 0
    _call("fbanet_faf_gate_bwd_sm100", p, nbytes=4 * feat.numel() * feat.element_size())
    return dfeat, dwsum


def faf_weight_grads(dwsum: torch.Tensor, cout: int):
    """Gradient of the two FAF embedding convolutions from the gate's ``dwsum [9,C]``: ``temporal_attn1.weight [cout,C,3,3]`` receives
    ``dwsum`` in every output channel (``wsum`` is the sum over them); ``temporal_attn0.weight`` and both biases receive zero --
    ``emb_ref`` and the biases cancel in ``aff[f] - aff[0]`` (``federated_affinity_fusion.py:84-99``)."""
    Cc = dwsum.shape[1]
    return dwsum.t().reshape(1, Cc, 3, 3).expand(cout, Cc, 3, 3)


def drop_path_add(x: torch.Tensor, scale: torch.Tensor, skip: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[b] = skip[b] + scale[b] * x[b]`` (``layers/drop_path.py:39-63`` "global" mode under ``jax.vmap``: one draw per burst;
    the residuals of ``layers/fba_net.py:245,248``); ``skip=None``: the branch's backward ``dx[b] = scale[b] * dy[b]``.
    ``scale``: fp32 ``[B]`` from :func:`fbanet_b200.train.drop_path_scales`."""
    assert x.is_cuda and x.is_contiguous() and scale.dtype == torch.float32 and scale.is_contiguous() and scale.numel() == x.shape[0]
    if skip is not None:
        assert skip.shape == x.shape and skip.dtype == x.dtype and skip.is_contiguous()
    out = torch.empty_like(x) if out is None else out
    assert out.shape == x.shape and out.dtype == x.dtype and out.is_contiguous()
    p = L.DropPathParams()
    p.x, p.out, p.scale = x.data_ptr(), out.data_ptr(), scale.data_ptr()
    p.skip = skip.data_ptr() if skip is not None else None
    p.per_burst, p.dtype, p.B = x.numel() // x.shape[0], _DT[x.dtype], x.shape[0]
    _call("fbanet_drop_path_add_sm100", p, nbytes=(3 if skip is not None else 2) * x.numel() * x.element_size())
    return out


def _band_params(bands, row0, tiles, T, Cc, H, W, psize, overlap, tile_begin, tile_end, scale):
    assert 1 <= len(bands) <= L.MAX_BANDS and len(row0) == len(bands) + 1 and row0[0] == 0 and row0[-1] == H
    p = L.TileBandParams()
    for k, b in enumerate(bands):
        p.band[k] = int(b)
        p.row0[k] = int(row0[k])
    p.row0[len(bands)] = H
    p.nbands, p.tiles = len(bands), tiles.data_ptr()
    p.T, p.C, p.H, p.W, p.psize, p.overlap, p.tile_begin, p.tile_end, p.scale = T, Cc, H, W, psize, overlap, tile_begin, tile_end, scale
    return p


def tile_divide_banded(bands, row0, T: int, Cc: int, H: int, W: int, psize: int, overlap: int, tile_begin: int, tile_end: int,
                       device) -> torch.Tensor:
    """Row-band form of :func:`tile_divide`: ``bands[k]`` is the device address (possibly peer memory) of the fp32 band
    ``[T,C,row0[k+1]-row0[k],W]``.  Returns the local tiles ``[n,T,C,ts,ts]``."""
    ts = psize + 2 * overlap
    out = torch.empty((tile_end - tile_begin, T, Cc, ts, ts), device=device, dtype=torch.float32)
    _call("fbanet_tile_divide_banded_sm100", _band_params(bands, row0, out, T, Cc, H, W, psize, overlap, tile_begin, tile_end, 1))
    return out


def tile_merge_banded(tiles: torch.Tensor, bands, row0, H: int, W: int, psize: int, overlap: int, scale: int, tile_begin: int,
                      tile_end: int) -> None:
    """Row-band form of :func:`tile_merge`: the x``scale`` centre of every tile is written into the output band that owns its rows
    (``bands[k]``: address of fp32 ``[C, scale*(row0[k+1]-row0[k]), scale*W]``, possibly peer memory)."""
    assert tiles.is_cuda and tiles.dtype == torch.float32 and tiles.is_contiguous() and tiles.shape[0] == tile_end - tile_begin
    _call("fbanet_tile_merge_banded_sm100", _band_params(bands, row0, tiles, 1, tiles.shape[1], H, W, psize, overlap, tile_begin, tile_end, scale))


def tile_divide(burst: torch.Tensor, psize: int, overlap: int, tile_begin: int = 0, tile_end: Optional[int] = None) -> torch.Tensor:
    """``[T,C,H,W]`` fp32 -> ``[tiles,T,C,psize+2ov,psize+2ov]`` (reflect padded, tile index row-major)."""
    assert burst.is_cuda and burst.dtype == torch.float32 and burst.is_contiguous() and burst.dim() == 4
    T, Cc, H, W = burst.shape
    nh, nw = -(-H // psize), -(-W // psize)
    tile_end = nh * nw if tile_end is None else tile_end
    ts = psize + 2 * overlap
    out = torch.empty((tile_end - tile_begin, T, Cc, ts, ts), device=burst.device, dtype=torch.float32)
    p = L.TileParams()
    p.src, p.dst = burst.data_ptr(), out.data_ptr()
    p.T, p.C, p.H, p.W, p.psize, p.overlap, p.tile_begin, p.tile_end, p.scale = T, Cc, H, W, psize, overlap, tile_begin, tile_end, 1
    _call("fbanet_tile_divide_sm100", p)
    return out


def tile_merge(tiles: torch.Tensor, out: torch.Tensor, H: int, W: int, psize: int, overlap: int, scale: int,
               tile_begin: int = 0, tile_end: Optional[int] = None) -> torch.Tensor:
    """SR tiles ``[n,C,scale*(psize+2ov)]^2`` -> centre-cropped stitch into ``out [C, scale*H, scale*W]``."""
    assert tiles.is_cuda and tiles.dtype == torch.float32 and tiles.is_contiguous() and out.is_contiguous()
    nh, nw = -(-H // psize), -(-W // psize)
    tile_end = nh * nw if tile_end is None else tile_end
    assert tiles.shape[0] == tile_end - tile_begin and out.shape == (tiles.shape[1], scale * H, scale * W)
    p = L.TileParams()
    p.src, p.dst = tiles.data_ptr(), out.data_ptr()
    p.T, p.C, p.H, p.W, p.psize, p.overlap, p.tile_begin, p.tile_end, p.scale = 1, tiles.shape[1], H, W, psize, overlap, tile_begin, tile_end, scale
    _call("fbanet_tile_merge_sm100", p)
    return out
