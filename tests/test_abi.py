"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/fbanet_b200.h
declares, and the ctypes structs match the compiled layout.  No kernels are launched."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_abi_matches(lib):
    from fbanet_b200 import _lib
    assert lib.fbanet_abi_version() == _lib.ABI_VERSION
    for name, st in _lib.STRUCTS.items():
        assert lib.fbanet_abi_sizeof(name.encode()) == ctypes.sizeof(st), name
    assert lib.fbanet_abi_sizeof(b"nope") == -1


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "fbanet_b200.h")).read()
    decl = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(fbanet_\w+)\s*\(", hdr, flags=re.M))
    from fbanet_b200 import _lib
    assert decl == set(_lib.OPS) | set(_lib.MISC_SYMBOLS), decl ^ (set(_lib.OPS) | set(_lib.MISC_SYMBOLS))
    for s in decl:
        assert hasattr(lib, s), s


def test_bad_params_are_rejected_without_a_gpu(lib):
    from fbanet_b200 import _lib
    p = _lib.ConvParams()
    assert lib.fbanet_conv_gemm_sm100(ctypes.byref(p), None) == -1  # null weight/out: bad shape, nothing launched
    a = _lib.AttnParams()
    assert lib.fbanet_window_attention_sm100(ctypes.byref(a), None) == -1
    w = _lib.WarpParams()
    assert lib.fbanet_warp_sm100(ctypes.byref(w), None) == -1
    # the front-end / multi-GPU entry points added for SURVEY 8e / 8f-4 refuse empty or inconsistent requests the same way
    for fn, st in (("fbanet_flow_warp_sm100", _lib.FlowWarpParams), ("fbanet_ecc_prepare_sm100", _lib.EccPrepareParams),
                   ("fbanet_ecc_homography_sm100", _lib.EccParams), ("fbanet_tile_divide_banded_sm100", _lib.TileBandParams),
                   ("fbanet_tile_merge_banded_sm100", _lib.TileBandParams)):
        assert getattr(lib, fn)(ctypes.byref(st()), None) == -1, fn
    b = _lib.TileBandParams()
    b.tiles, b.nbands, b.H, b.W, b.T, b.C, b.psize, b.overlap, b.tile_end, b.scale = 1, 2, 50, 70, 1, 1, 20, 10, 1, 1
    b.band[0], b.band[1] = 1, 1
    b.row0[0], b.row0[1], b.row0[2] = 0, 25, 49          # bands must tile [0, H) exactly
    assert lib.fbanet_tile_divide_banded_sm100(ctypes.byref(b), None) == -1
    b.row0[2], b.row0[1] = 50, 0                          # empty band
    assert lib.fbanet_tile_divide_banded_sm100(ctypes.byref(b), None) == -1
    # backward bricks of the training step: empty requests, inconsistent output sizes, too-wide LayerNorm rows
    for fn, st in (("fbanet_wgrad_sm100", _lib.WgradParams), ("fbanet_layernorm_bwd_sm100", _lib.LayerNormBwdParams),
                   ("fbanet_act_bwd_sm100", _lib.ActBwdParams)):
        assert getattr(lib, fn)(ctypes.byref(st()), None) == -1, fn
    g = _lib.WgradParams()
    g.x, g.dy, g.dw, g.partial = 1, 1, 1, 1
    g.N, g.H, g.W, g.Cin, g.Cout, g.KH, g.KW, g.stride, g.pad, g.splits, g.x_ld, g.dy_ld = 1, 8, 8, 4, 4, 3, 3, 1, 1, 1, 4, 4
    g.Ho, g.Wo = 8, 7                                     # Wo must be (W + 2 pad - KW) / stride + 1
    assert lib.fbanet_wgrad_sm100(ctypes.byref(g), None) == -1
    n = _lib.LayerNormBwdParams()
    n.x, n.dy, n.gamma, n.dx, n.partial, n.rows, n.C, n.eps = 1, 1, 1, 1, 1, 4, 512, 1e-5   # C > 256
    assert lib.fbanet_layernorm_bwd_sm100(ctypes.byref(n), None) == -1
    assert lib.fbanet_layernorm_bwd_blocks(0) == -1 and lib.fbanet_act_bwd_blocks(0) == -1
    # second set of backward bricks (depthwise 3x3, window attention, FAF gate, DropPath residual)
    for fn, st in (("fbanet_dwconv3x3_bwd_sm100", _lib.DwconvBwdParams), ("fbanet_window_attention_bwd_sm100", _lib.AttnBwdParams),
                   ("fbanet_faf_gate_bwd_sm100", _lib.FafGateBwdParams), ("fbanet_drop_path_add_sm100", _lib.DropPathParams)):
        assert getattr(lib, fn)(ctypes.byref(st()), None) == -1, fn
    f = _lib.ActFwdParams()
    assert lib.fbanet_act_fwd_sm100(ctypes.byref(f), None) == -1
    f.x, f.y, f.n, f.act = 1, 1, 8, 2                                                             # PReLU without its slope
    assert lib.fbanet_act_fwd_sm100(ctypes.byref(f), None) == -1
    d = _lib.DwconvBwdParams()
    d.x, d.dy, d.weight, d.partial, d.N, d.H, d.W, d.C, d.dtype = 1, 1, 1, 1, 1, 4, 4, 8, 0      # no output requested
    assert lib.fbanet_dwconv3x3_bwd_sm100(ctypes.byref(d), None) == -1
    t = _lib.AttnBwdParams()
    t.qkv, t.dout, t.dqkv, t.bias_table, t.dbias = 1, 1, 1, 1, 1                                  # table gradient without its workspace
    t.B, t.H, t.W, t.C, t.heads, t.win, t.qkv_ld, t.dout_ld, t.dqkv_ld = 1, 10, 10, 64, 1, 10, 192, 64, 192
    assert lib.fbanet_window_attention_bwd_sm100(ctypes.byref(t), None) == -1
    t.dbias, t.H = None, 12                                                                        # H not a multiple of win
    assert lib.fbanet_window_attention_bwd_sm100(ctypes.byref(t), None) == -1
    assert lib.fbanet_attn_bwd_partial_floats(2, 20, 20, 4, 10) == 2 * 4 * 4 * 361 and lib.fbanet_attn_bwd_partial_floats(1, 12, 10, 1, 10) == -1
    assert lib.fbanet_dwconv_bwd_blocks(0) == -1 and lib.fbanet_faf_gate_bwd_blocks(0) == -1 and lib.fbanet_dwconv_bwd_blocks(9) == 3
    # round-2 options that exist on the tensor-core path only are refused, not silently dropped, when a problem falls back to the
    # CUDA-core kernel (fp32 here): LayerNorm inside the GEMM, fp16 store; the fused homography warp of the head conv likewise
    c = _lib.ConvParams()
    c.nsrc, c.weight, c.out, c.dtype = 1, 1, 1, 0
    c.src[0].ptr, c.src[0].C, c.src[0].ld, c.src[0].img_stride = 1, 64, 64, 64 * 64
    c.N, c.H, c.W, c.KH, c.KW, c.stride, c.pad, c.Ho, c.Wo, c.Cout, c.Cout_store, c.out_ld, c.out_img_stride = 1, 8, 8, 1, 1, 1, 0, 8, 8, 64, 64, 64, 64 * 64
    c.ln_gamma, c.ln_beta = 1, 1
    assert lib.fbanet_conv_gemm_sm100(ctypes.byref(c), None) == -5
    c.ln_gamma, c.ln_beta, c.store_f16 = None, None, 1
    assert lib.fbanet_conv_gemm_sm100(ctypes.byref(c), None) == -5
    hc = _lib.HeadConvParams()
    hc.src, hc.dst, hc.weight, hc.bias, hc.dtype, hc.frames, hc.C, hc.H, hc.W, hc.Cout = 256, 256, 256, 256, 0, 2, 3, 8, 8, 64
    hc.M, hc.frames_per_burst = 256, 2
    assert lib.fbanet_head_conv_sm100(ctypes.byref(hc), None) == -5          # fp32 destination: no fused warp
    hc.frames_per_burst = 0
    assert lib.fbanet_head_conv_sm100(ctypes.byref(hc), None) == -1
    e = _lib.EccParams()
    e.planes, e.warp, e.frames, e.frames_per_burst, e.H, e.W, e.max_iters = 1, 1, 7, 2, 16, 16, 10   # 7 frames are not whole bursts of 2
    assert lib.fbanet_ecc_homography_sm100(ctypes.byref(e), None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from fbanet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfbanet_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
