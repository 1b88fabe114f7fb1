"""Whole-model parity of the CUDA BaseModel against the CPU oracle (random init, identical weights).

Bars (BASELINE.json north_star): fp32 max-abs <= 1e-3; bf16 PSNR delta <= 0.01 dB.  The PSNR delta is
measured against a synthetic ground truth (bilinear x4 of the base frame + fixed noise), i.e.
|PSNR(bf16 out, gt) - PSNR(oracle out, gt)|; the direct bf16-vs-oracle PSNR is asserted too."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(num_frames=4, img_size=40, in_channels=3, embed_dim=32, window_length=10)
FULL = dict(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10)
RAW = dict(num_frames=14, img_size=80, in_channels=4, embed_dim=64, window_length=10)


def _pair(cfg, dtype, cuda, seed=0):
    from fbanet_b200 import BaseModel
    from oracle.fbanet_oracle import build_oracle
    o = build_oracle(seed, **cfg)
    m = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype=dtype)
    m.load_state_dict(o.state_dict())
    return o, m.to(cuda).eval()


def _burst(cfg, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, cfg["num_frames"], cfg["in_channels"], cfg["img_size"], cfg["img_size"], generator=g)


def _to_oracle_layout(name, t, ref):
    t = t.float().cpu()
    if name in ("head", "body"):          # [B,F,H,W,E] -> [B,F,E,H,W]
        return t.permute(0, 1, 4, 2, 3)
    if name in ("faf.gate", "out"):
        return t
    if ref.dim() == 3:                     # tokens [B,T,C]
        return t.reshape(t.shape[0], -1, t.shape[-1])
    return t.permute(0, 3, 1, 2)           # image [B,C,H,W]


def test_small_model_every_stage_fp32(cuda):
    o, m = _pair(SMALL, "fp32", cuda)
    x = _burst(SMALL, 2)
    with torch.no_grad():
        ref = o.forward_stages(x)
    st = {}
    m.forward_stages(x.to(cuda), st)
    worst = {}
    for k, r in ref.items():
        if k not in st:
            continue
        got = _to_oracle_layout(k, st[k], r)
        assert got.shape == r.shape, (k, got.shape, r.shape)
        worst[k] = (got - r).abs().max().item()
    print(worst)
    assert len(worst) >= 30
    bad = {k: v for k, v in worst.items() if v > 2e-4}
    assert not bad, bad


def test_small_model_bf16_close(cuda):
    from oracle.fbanet_oracle import psnr
    o, m = _pair(SMALL, "bf16", cuda)
    x = _burst(SMALL, 2)
    with torch.no_grad():
        ref = o(x)
    got = m(x.to(cuda)).cpu()
    assert psnr(got, ref) > 40.0, psnr(got, ref)


def test_input_shape_assert(cuda):
    _, m = _pair(SMALL, "fp32", cuda)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 4, 3, 32, 32, device=cuda))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 3, 40, 40))  # CPU tensor: no CPU fallback


@pytest.mark.parametrize("cfg", [FULL, RAW], ids=["rgb160", "raw80"])
def test_full_config_fp32_within_1e3(cuda, cfg):
    """BASELINE configs[0] / configs[2] shapes, batch 1: max abs error <= 1e-3 in fp32."""
    o, m = _pair(cfg, "fp32", cuda)
    x = _burst(cfg, 1)
    with torch.no_grad():
        ref = o(x)
    got = m(x.to(cuda)).cpu()
    err = (got - ref).abs().max().item()
    print("max abs err", err)
    assert got.shape == ref.shape == (1, cfg["in_channels"], 4 * cfg["img_size"], 4 * cfg["img_size"])
    assert err <= 1e-3, err


def _psnr_delta(cuda, seed, fold_ln=False):
    from oracle.fbanet_oracle import psnr
    o, m = _pair(FULL, "bf16", cuda, seed=seed)
    m.fold_ln = fold_ln
    x = _burst(FULL, 1, seed=seed)
    with torch.no_grad():
        ref = o(x)
    got = m(x.to(cuda)).cpu()
    g = torch.Generator().manual_seed(7)
    gt = (torch.nn.functional.interpolate(x[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
          + 0.05 * torch.randn(ref.shape, generator=g)).clamp(0, 1)
    delta = abs(psnr(got.clamp(0, 1), gt) - psnr(ref.clamp(0, 1), gt))
    direct = psnr(got, ref)
    print("seed", seed, "fold_ln", fold_ln, "psnr delta", delta, "direct psnr", direct)
    return delta, direct


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5, 6, 7])
def test_full_config_bf16_psnr_delta(cuda, seed):
    """north-star tolerance for the bf16 path: |PSNR(bf16, gt) - PSNR(oracle, gt)| <= 0.01 dB on the full cfg2 shape, for eight
    independent weight / burst draws (the delta is a projection of the fixed weight-rounding perturbation onto the image).
    Fails at 0.01 dB, warns from 0.008 dB."""
    import warnings
    delta, direct = _psnr_delta(cuda, seed)
    assert delta <= 0.01, delta
    assert direct >= 40.0, direct
    if delta > 0.008:
        warnings.warn(f"bf16 PSNR delta {delta:.4f} dB for seed {seed} is within 20 % of the 0.01 dB tolerance")


def test_full_config_bf16_batch_invariance_and_graph_replay(cuda):
    """The benchmarked path itself (bf16, embed 64, 160x160: every contraction on the tcgen05 kernels): a batch of 4 equals four
    single-burst forwards BIT FOR BIT (bursts are independent units, jax.vmap(model), train.py:35 -- no kernel may let the tile
    schedule or a neighbouring burst leak into a result), and the CUDA-graph replay of `infer_host` (whole batch and 2-burst
    chunks) equals the eager forward."""
    _, m = _pair(FULL, "bf16", cuda, seed=1)
    x = _burst(FULL, 4, seed=4)
    xd = x.to(cuda)
    full = m(xd).cpu()
    for i in range(4):
        assert torch.equal(m(xd[i:i + 1]).cpu(), full[i:i + 1]), i
    assert m.host_graphs
    for chunk in (4, 2):
        for _ in range(2):
            assert torch.equal(m.infer_host(x, chunk=chunk), full), chunk


def test_full_config_bf16_layernorm_fold(cuda):
    """optional `fold_ln` mode (LayerNorm folded into the qkv / fc1 GEMMs): same direct PSNR against the oracle; its PSNR delta
    is 0.0105 / 0.0007 / 0.0016 dB over the three draws -- mean within the 0.01 dB tolerance, the first draw 5 % over it, which
    is why the mode is off by default."""
    res = [_psnr_delta(cuda, seed, fold_ln=True) for seed in (0, 1, 2)]
    assert sum(d for d, _ in res) / 3 <= 0.01, res
    assert max(d for d, _ in res) <= 0.02, res
    assert min(p for _, p in res) >= 60.0, res


def test_batch_invariance_and_host_api(cuda):
    """bursts are independent units (jax.vmap(model), train.py:35): batch of 3 == three singles; and the
    host-buffer entry point returns the same image."""
    _, m = _pair(SMALL, "fp32", cuda)
    x = _burst(SMALL, 3, seed=3)
    full = m(x.to(cuda)).cpu()
    for i in range(3):
        one = m(x[i:i + 1].to(cuda)).cpu()
        assert torch.equal(one, full[i:i + 1])
    host = m.infer_host(x)
    assert torch.equal(host, full)
    # chunked, three-stream pipeline (H2D / forward / D2H overlapped): any chunk size gives the same image, repeatedly
    for chunk in (1, 2, 3):
        for _ in range(2):
            assert torch.equal(m.infer_host(x, chunk=chunk), full)


def test_host_api_streaming(cuda):
    """Asynchronous host calls (`infer_host(wait=False)` handles, `infer_host_stream` over a loader-like iterable): uploads,
    forwards and downloads of consecutive batches overlap, the buffer slots alternate across calls -- and every batch still
    gets exactly its own image, in order, eager and graph-replayed, with ragged batch sizes in the stream."""
    _, m = _pair(SMALL, "fp32", cuda)
    batches = [_burst(SMALL, n, seed=20 + i) for i, n in enumerate((2, 2, 2, 3, 2, 1, 2))]
    want = [m(b.to(cuda)).cpu() for b in batches]
    for graphs in (True, False):
        m.host_graphs = graphs
        # handles: everything queued before anything is waited for
        outs = [torch.empty_like(w).pin_memory() for w in want]
        hs = [m.infer_host(b.pin_memory(), o, wait=False) for b, o in zip(batches, outs)]
        for h, w in zip(reversed(hs), reversed(want)):
            got = h.wait()
            assert h.done() and torch.equal(got, w)
        # generator, own rotating buffers: compare as they are yielded
        for depth in (1, 2, 3):
            n = 0
            for got, w in zip(m.infer_host_stream(batches, depth=depth), want):
                assert torch.equal(got, w), (graphs, depth, n)
                n += 1
            assert n == len(want)
        # a blocking call between streamed ones, and narrow output
        h = m.infer_host(batches[0], wait=False)
        assert torch.equal(m.infer_host(batches[3]), want[3])
        assert torch.equal(h.wait(), want[0])
        # (graphs of several sizes / dtypes share one activation pool: a result buffer inside it would be overwritten by the
        # next replay while its download is still running -- seen as torn uint8 images before the buffers moved out of the pool)
        for rep in range(3):
            for depth in (2, 3):
                u8 = [g.clone() for g in m.infer_host_stream(batches, out_dtype=torch.uint8, depth=depth)]
                for i, (got, w) in enumerate(zip(u8, want)):
                    assert torch.equal(got, torch.clamp(w, 0, 1).mul(255).byte()), (graphs, rep, depth, i)
    assert m.infer_host(batches[0][:0], wait=False).wait().shape[0] == 0


def test_host_api_narrow_io(cuda):
    """`infer_host` with the reference's 8-bit data path: uint8 frames are normalised on the device exactly as `x.astype(float32) /
    255.0` (train.py:82-83), a uint8 result is `clamp(SR, 0, 1).mul(255).byte()` (test_in_any_resolution.py:93 + torchvision
    ToPILImage: truncation), an fp16 result is the rounded fp32 one; eager and graph-replay paths agree."""
    _, m = _pair(SMALL, "fp32", cuda)
    g = torch.Generator().manual_seed(9)
    xu = torch.randint(0, 256, (3, SMALL["num_frames"], SMALL["in_channels"], SMALL["img_size"], SMALL["img_size"]), generator=g, dtype=torch.uint8)
    xf = xu.float() / 255.0
    ref = m(xf.to(cuda)).cpu()
    assert torch.equal(m.infer_host(xu), ref)
    assert torch.equal(m.infer_host(xf, out_dtype=torch.float16), ref.half())
    want = torch.clamp(ref, 0, 1).mul(255).byte()
    for chunk in (3, 2):
        got = m.infer_host(xu, out_dtype=torch.uint8, chunk=chunk)
        assert got.dtype == torch.uint8 and torch.equal(got, want)
    assert 0 < want.float().mean().item() < 255            # not saturated: the comparison means something
    m.host_graphs = False
    assert torch.equal(m.infer_host(xu, out_dtype=torch.uint8), want)


def test_fhwc_adaptor(cuda):
    _, m = _pair(SMALL, "fp32", cuda)
    x = _burst(SMALL, 1)
    a = m(x.to(cuda))[0].permute(1, 2, 0)
    b = m.forward_fhwc(x[0].permute(0, 2, 3, 1).to(cuda))
    assert torch.equal(a, b)


def test_golden_fixtures_on_gpu(cuda):
    """The committed golden vectors (tests/golden) through the CUDA path."""
    import os
    import numpy as np
    from fbanet_b200 import BaseModel, ops
    from oracle.fbanet_oracle import build_oracle
    gold = os.path.join(os.path.dirname(__file__), "golden")
    g = torch.load(os.path.join(gold, "small_model.pt"))
    m = BaseModel(**g["cfg"], token_projection="linear", token_mlp="leff", dtype="fp32")
    m.load_state_dict(build_oracle(g["seed"], **g["cfg"]).state_dict())
    got = m.to(cuda)(g["x"].to(cuda)).cpu()
    assert (got - g["out"]).abs().max().item() <= 1e-4
    w = np.load(os.path.join(gold, "warp.npz"))
    out = ops.warp_burst(torch.from_numpy(w["burst"])[None].to(cuda), torch.from_numpy(w["M"])[None], layout="BTHWC").cpu().numpy()[0]
    assert np.abs(out - w["out"]).max() < 2e-6


def test_tiled_full_resolution_matches_reference_driver(cuda):
    """BASELINE config 4 in miniature: reflect-pad tiling + per-tile forward + centre-crop stitch against the
    restated test_in_any_resolution.py loop (oracle model, oracle tiling), incl. a 2-rank tile shard."""
    from fbanet_b200.tiling import infer_full_resolution
    from oracle.fbanet_oracle import tensor_divide_burst, tensor_merge
    cfg = dict(num_frames=3, img_size=40, in_channels=3, embed_dim=32, window_length=10)
    o, m = _pair(cfg, "fp32", cuda)
    x = torch.rand(1, 3, 3, 50, 70, generator=torch.Generator().manual_seed(5))
    tiles = tensor_divide_burst(x, 20, 10)
    with torch.no_grad():
        sr = torch.cat([o(tiles[i:i + 1]) for i in range(tiles.shape[0])], 0)
    ref = tensor_merge(sr, (200, 280), psize=80, overlap=40)
    got = infer_full_resolution(m, x.to(cuda), psize=20, overlap=10, tile_batch=5).cpu()
    assert got.shape == ref.shape == (1, 3, 200, 280)
    assert (got - ref).abs().max().item() <= 1e-3
    # tile sharding over 2 ranks: the two partial canvases add up to the full image
    out = torch.zeros(3, 200, 280, device=cuda)
    for r in range(2):
        infer_full_resolution(m, x.to(cuda), psize=20, overlap=10, tile_batch=4, rank=r, world=2, out=out)
    assert (out.cpu()[None] - ref).abs().max().item() <= 1e-3


def test_raw_config_with_warp_front_end_bf16(cuda):
    """BASELINE config 3 shape: 14x4x80x80 packed-Bayer burst, homography warp + FAF + SR, bf16 path."""
    import numpy as np
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import psnr, warp_burst as warp_ref
    o, m = _pair(RAW, "bf16", cuda)
    x = _burst(RAW, 1, seed=2)
    rng = np.random.default_rng(1)
    M = np.tile(np.eye(3), (1, 14, 1, 1))
    M[..., :2, :2] += rng.uniform(-0.01, 0.01, (1, 14, 2, 2))
    M[..., :2, 2] += rng.uniform(-4, 4, (1, 14, 2))
    M[..., 2, :2] += rng.uniform(-1e-5, 1e-5, (1, 14, 2))
    xw = ops.warp_burst(x.to(cuda), torch.from_numpy(M))
    ref_w = torch.from_numpy(warp_ref(x[0].permute(0, 2, 3, 1).numpy(), M[0])).permute(0, 3, 1, 2).float()[None]
    assert (xw.cpu() - ref_w).abs().max().item() < 2e-6
    with torch.no_grad():
        ref = o(ref_w)
    got = m(xw).cpu()
    assert got.shape == (1, 4, 320, 320)
    assert psnr(got, ref) > 40.0


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_raw_config_batch64_perspective_warp_against_oracle(cuda, dtype):
    """BASELINE config 3 at its full batch: 64 packed-Bayer bursts (14x4x80x80), every frame warped by its own homography with a
    NON-ZERO perspective row (translation U(-4,4) px, affine U(-0.01,0.01), perspective U(-1e-5,1e-5)), warp + forward
    in one batch-64 call; one burst of each 32-burst host chunk is checked against the oracle (fp32: <= 1e-3 max abs and warp
    <= 2e-6; bf16: PSNR > 40 dB)."""
    import numpy as np
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import psnr, warp_burst as warp_ref
    o, m = _pair(RAW, dtype, cuda)
    B, T = 64, RAW["num_frames"]
    x = _burst(RAW, B, seed=3)
    g = torch.Generator().manual_seed(1)
    M = torch.eye(3, dtype=torch.float64).repeat(B, T, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 8 - 4
    M[:, 1:, :2, :2] += torch.rand(B, T - 1, 2, 2, generator=g, dtype=torch.float64) * 0.02 - 0.01
    M[:, 1:, 2, :2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 2e-5 - 1e-5
    xw = ops.warp_burst(x.to(cuda), M.to(cuda))
    got = m(xw).cpu()
    assert got.shape == (B, 4, 320, 320)
    for i in (5, 37):
        ref_w = torch.from_numpy(np.asarray(warp_ref(x[i].permute(0, 2, 3, 1).numpy(), M[i].numpy()), np.float32)).permute(0, 3, 1, 2)[None].contiguous()
        assert (xw[i:i + 1].cpu() - ref_w).abs().max().item() < 2e-6
        with torch.no_grad():
            ref = o(ref_w)
        if dtype == "fp32":
            assert (got[i:i + 1] - ref).abs().max().item() <= 1e-3, (i, (got[i:i + 1] - ref).abs().max().item())
        else:
            assert psnr(got[i:i + 1], ref) > 40.0, (i, psnr(got[i:i + 1], ref))


def test_forward_unaligned_registers_then_restores(cuda):
    """8f-4 + the forward in one call: a burst whose frames are shifted copies of the base frame must, after the on-device ECC +
    warp, give (nearly) the SR image of the perfectly aligned burst -- and be far from the SR image of the unaligned one."""
    from fbanet_b200 import ops
    from oracle.fbanet_oracle import psnr
    _, m = _pair(SMALL, "fp32", cuda)
    S, T, C = SMALL["img_size"], SMALL["num_frames"], SMALL["in_channels"]
    g = torch.Generator().manual_seed(11)
    base = torch.rand(1, 1, C, S + 16, S + 16, generator=g)
    k = torch.ones(C, 1, 5, 5) / 25.0
    for _ in range(3):
        base = torch.nn.functional.conv2d(torch.nn.functional.pad(base[:, 0], (2, 2, 2, 2), mode="reflect"), k, groups=C)[:, None]
    base = base[..., 8:-8, 8:-8].contiguous()
    base = (base - base.amin()) / (base.amax() - base.amin())
    aligned = base.expand(1, T, C, S, S).contiguous().to(cuda)
    M = torch.eye(3, dtype=torch.float64).repeat(1, T, 1, 1)
    M[0, 1:, :2, 2] = torch.rand(T - 1, 2, generator=g, dtype=torch.float64) * 3 - 1.5
    shifted = ops.warp_burst(aligned, M)
    shifted[:, 0] = aligned[:, 0]
    y_ref = m(aligned).cpu()
    y_raw = m(shifted).cpu()
    y_reg, (Mh, rho, iters) = m.forward_unaligned(shifted, return_alignment=True)
    assert (iters[:, 1:] > 0).all() and (rho[:, 1:] > 0.9).all()      # 40 x 40 frames: the warp border costs some correlation
    reg = ops.warp_burst(shifted, Mh)
    fi = (slice(None), slice(1, None), slice(None), slice(6, -6), slice(6, -6))   # the warps leave a zero border
    assert psnr(reg[fi].cpu(), aligned[fi].cpu()) > psnr(shifted[fi].cpu(), aligned[fi].cpu()) + 6.0      # the frames are registered
    inner = (slice(None), slice(None), slice(32, -32), slice(32, -32))
    # a random-init network leans mostly on the base frame, so the SR gain is modest (measured +2.1 dB), but it must be a gain
    assert psnr(y_reg.cpu()[inner], y_ref[inner]) > psnr(y_raw[inner], y_ref[inner]) + 1.0


def test_full_resolution_row_bands_two_gpus(cuda):
    """cfg4 / SURVEY 8e on real peer memory: two ranks, the burst sharded by row bands in symmetric memory, halos read over NVLink
    inside the tile-gather kernel -- the stitched image must be bit-identical to the single-GPU replicated driver.  Needs two
    GPUs (skipped on the one-GPU box; run with `gpurun --gpus 2`)."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(root, "tools", "run_fullres.py"), "--H", "200", "--W", "330", "--steps", "1", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["bit_identical"] and res["n_gpus"] == 2 and res["halo_bytes_read_from_peers_per_image"] > 0


def test_empty_batch(cuda):
    """An empty batch is a valid input (a shard can be empty when bursts < ranks): empty output, no launch, on both entry points."""
    _, m = _pair(SMALL, "bf16", cuda)
    S, T, C = SMALL["img_size"], SMALL["num_frames"], SMALL["in_channels"]
    y = m(torch.empty(0, T, C, S, S, device=cuda))
    assert y.shape == (0, C, 4 * S, 4 * S) and y.dtype == torch.float32
    assert m.infer_host(torch.empty(0, T, C, S, S)).shape == (0, C, 4 * S, 4 * S)
    with pytest.raises(AssertionError):
        m(torch.empty(1, T, C, S + 1, S, device=cuda))      # wrong frame size: the reference's shape assert (models/fba_net.py:244)


@pytest.mark.parametrize("pre,dtype", [("faf_gpu", "fp32"), ("faf_gpu64", "fp32"), ("faf_gpu64", "bf16")])
def test_faf_against_reference_executed_fixture(cuda, pre, dtype):
    """CUDA FAF (gate kernel + fuse GEMM + hourglass convs) against the output of the REFERENCE'S OWN `FAFBlock` code
    (tests/golden/layers_reference.npz `faf_gpu*/`, produced by tests/golden/make_golden_layers.py executing
    blocks/federated_affinity_fusion.py:18-182).  The oracle module is only the carrier of the layout-converted weights here;
    the values compared are the reference's.  bf16 with 64 channels runs the tcgen05 implicit-GEMM path."""
    import numpy as np
    from fbanet_b200 import BaseModel
    from oracle.fbanet_oracle import psnr
    from test_oracle_reference_layers import GOLD, faf_gpu_case

    z = np.load(GOLD)
    carrier, feat, gate, y = faf_gpu_case({k: z[k] for k in z.files if k.startswith(pre + "/")}, pre)
    nf, frames, side = feat.shape[2], feat.shape[1], feat.shape[3]
    m = BaseModel(num_frames=frames, img_size=side, in_channels=3, embed_dim=nf, window_length=10, token_projection="linear",
                  token_mlp="leff", dtype=dtype)
    m.fusion.load_state_dict(carrier.state_dict())
    m = m.to(cuda).eval()
    x = feat.permute(0, 1, 3, 4, 2).contiguous().to(cuda).to(m.compute_dtype)  # [1,F,H,W,E]
    out, _, g = m._faf(m.packed(), x)
    torch.cuda.synchronize()
    out, g = out.float().cpu()[0].numpy(), g.float().cpu()[0].numpy()
    assert out.shape == y.shape and g.shape == gate.shape
    if dtype == "fp32":
        assert np.abs(g - gate).max() <= 1e-4, np.abs(g - gate).max()
        assert np.abs(out - y).max() <= 1e-3, np.abs(out - y).max()          # north_star: fp32 max abs error <= 1e-3
    else:
        assert np.abs(g - gate).max() <= 2e-2, np.abs(g - gate).max()
        span = float(y.max() - y.min())
        assert psnr(torch.from_numpy(out), torch.from_numpy(y), max_val=span) > 40.0


def _ref_fixture(prefixes):
    import numpy as np
    from test_oracle_reference_layers import GOLD

    z = np.load(GOLD)
    return {k: z[k] for k in z.files if k.split("/")[0] in prefixes}


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_window_attention_against_reference_executed_fixture(cuda, dtype):
    """qkv GEMM -> window attention kernel -> proj GEMM of the first encoder stage (dim 64, window 10, heads 1) against the output of
    the REFERENCE'S OWN `WindowAttentionLayer` code (layers/window_attention.py:159-248 through `jax.vmap` over four windows,
    `attn_gpu/*` of tests/golden/layers_reference.npz).  bf16 runs the tensor-core attention and the tcgen05 GEMMs."""
    import numpy as np
    from fbanet_b200 import BaseModel, ops
    from oracle.fbanet_oracle import psnr, window_partition, window_reverse
    from test_oracle_reference_layers import attn_gpu_case

    carrier, xw, y = attn_gpu_case(_ref_fixture({"attn_gpu"}))
    nwin, N, dim = xw.shape
    win, side = int(round(N ** 0.5)), int(round((nwin * N) ** 0.5))
    m = BaseModel(num_frames=2, img_size=side, in_channels=3, embed_dim=dim, window_length=win, token_projection="linear",
                  token_mlp="leff", dtype=dtype)
    ly = m.HG1_encoderlayer_0.blocks[0]
    assert (ly.heads, ly.win, ly.shift, ly.dim) == (1, win, 0, dim)
    ly.attn.load_state_dict(carrier.state_dict())
    m = m.to(cuda).eval()
    P, key = m.packed(), "HG1_encoderlayer_0.0"
    x = window_reverse(xw, win, 1, side, side).contiguous().to(cuda).to(m.compute_dtype)  # [1,H,W,d]
    qkv = m._lin(P, key + ".qkv", x)
    att = ops.window_attention(qkv.view(-1, 3 * dim), P[key + ".rpb"], 1, side, side, 1, win, 0, dim ** -0.5, impl=m.impl,
                               bias_expanded=P.get(key + ".rpbx"), q_prescaled=m._use_tc())
    out = m._lin(P, key + ".proj", att.view(1, side, side, dim))
    torch.cuda.synchronize()
    got = window_partition(out.float().cpu(), win).numpy()
    assert got.shape == y.shape
    if dtype == "fp32":
        assert np.abs(got - y).max() <= 1e-3, np.abs(got - y).max()
    else:
        assert psnr(torch.from_numpy(got), torch.from_numpy(y), max_val=float(y.max() - y.min())) > 40.0


def test_qkv_layout_against_reference_executed_fixture(cuda):
    """The fused q | k | v GEMM's column layout (head-major inside each third) against the reference's own
    `LinearProjectionLayer` split `"n (hd h c) -> hd h n c"` (layers/linear_projection.py:38-43) at dim 128, heads 2."""
    import numpy as np
    from fbanet_b200 import BaseModel
    from test_oracle_reference_layers import qkv_gpu_case

    carrier, x, q, k, v = qkv_gpu_case(_ref_fixture({"qkv_gpu"}))
    heads, n, dh = q.shape
    dim = heads * dh
    m = BaseModel(num_frames=2, img_size=20, in_channels=3, embed_dim=dim // 2, window_length=10, token_projection="linear",
                  token_mlp="leff", dtype="fp32")
    ly = m.HG1_encoderlayer_1.blocks[0]
    assert (ly.heads, ly.dim) == (heads, dim)
    ly.attn.qkv.load_state_dict(carrier.state_dict())
    m = m.to(cuda).eval()
    qkv = m._lin(m.packed(), "HG1_encoderlayer_1.0.qkv", x.view(1, 10, 10, dim).contiguous().to(cuda))
    torch.cuda.synchronize()
    got = qkv.float().cpu().view(n, 3, heads, dh).permute(1, 2, 0, 3).numpy()  # [3, h, n, dh]
    for g, r in zip(got, (q, k, v)):
        assert np.abs(g - r).max() <= 1e-4, np.abs(g - r).max()
