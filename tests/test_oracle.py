"""CPU checks that pin the oracle (oracle/fbanet_oracle.py) against independent closed forms and the
committed golden fixtures.  The reference ships no tests / golden vectors, so these are
the pins: numpy formulas, float64 restatements, cv2.warpPerspective in 1/32-px mode."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fbanet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_parameter_counts_match_survey():
    assert sum(p.numel() for p in O.OracleBaseModel().parameters()) == 19_217_237
    assert sum(p.numel() for p in O.OracleBaseModel(in_channels=4, img_size=80).parameters()) == 19_218_390


def test_relative_position_index_closed_form():
    w = 10
    idx = O.relative_position_index(w).numpy()
    ys, xs = np.divmod(np.arange(w * w), w)
    ref = (ys[:, None] - ys[None, :] + w - 1) * (2 * w - 1) + (xs[:, None] - xs[None, :] + w - 1)
    assert np.array_equal(idx, ref)
    assert idx.min() == 0 and idx.max() == 360 and len(np.unique(idx)) == 361  # Appendix A-2


def test_shift_mask_nine_regions_brute_force():
    H = W = 40
    win, s = 10, 5
    mask = O.shift_attn_mask(H, W, win, s).numpy()
    rid = lambda v, L: 0 if v < L - win else (1 if v < L - s else 2)
    assert mask.shape == (16, 100, 100)
    for wy, wx in ((0, 0), (3, 0), (0, 3), (3, 3), (1, 2)):
        ids = np.array([rid(wy * win + i // win, H) * 3 + rid(wx * win + i % win, W) for i in range(100)])
        ref = np.where(ids[None, :] != ids[:, None], -100.0, 0.0)
        assert np.array_equal(mask[wy * 4 + wx], ref)
    assert (mask[0] == 0).all() and (mask[15] != 0).any()


def test_window_partition_reverse_roundtrip_and_order():
    x = torch.arange(2 * 20 * 30 * 3, dtype=torch.float32).view(2, 20, 30, 3)
    w = O.window_partition(x, 10)
    assert w.shape == (12, 100, 3)
    assert torch.equal(w[4, 23], x[0, 10 + 2, 10 + 3])  # window (1,1), token (2,3)
    assert torch.equal(O.window_reverse(w, 10, 2, 20, 30), x)


def test_pixel_shuffle_order_is_torch():
    """Appendix A-18: out[c, 2y+i, 2x+j] = in[4c+2i+j, y, x]."""
    x = torch.rand(1, 8, 3, 5)
    y = F.pixel_shuffle(x, 2)
    for c in range(2):
        for i in range(2):
            for j in range(2):
                assert torch.equal(y[0, c, i::2, j::2], x[0, 4 * c + 2 * i + j])


def test_bilinear_base_formula():
    """Appendix A-19 / D: half-pixel centres, source coordinate clamped at 0, edge-clamped neighbour."""
    x = torch.rand(1, 1, 6, 7, dtype=torch.float64)
    y = F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)[0, 0].numpy()
    xs = x[0, 0].numpy()
    for Y, X in ((0, 0), (1, 2), (13, 9), (23, 27), (10, 0)):
        sy, sx = max((Y + 0.5) / 4 - 0.5, 0), max((X + 0.5) / 4 - 0.5, 0)
        y0, x0 = int(sy), int(sx)
        y1, x1 = min(y0 + 1, 5), min(x0 + 1, 6)
        ly, lx = sy - y0, sx - x0
        ref = (1 - ly) * ((1 - lx) * xs[y0, x0] + lx * xs[y0, x1]) + ly * ((1 - lx) * xs[y1, x0] + lx * xs[y1, x1])
        assert abs(y[Y, X] - ref) < 1e-12


def test_gelu_is_tanh_approximation():
    u = torch.linspace(-6, 6, 101, dtype=torch.float64)
    ref = 0.5 * u * (1 + torch.tanh(np.sqrt(2 / np.pi) * (u + 0.044715 * u ** 3)))
    assert torch.allclose(O.gelu_fn("tanh")(u), ref, atol=1e-12)
    assert not torch.allclose(O.gelu_fn("erf")(u), ref, atol=1e-5)


def test_warp_identity_and_integer_translation():
    img = np.random.default_rng(0).random((12, 15, 3))
    assert np.allclose(O.warp_frame(img, np.eye(3)), img)
    M = np.eye(3)
    M[0, 2], M[1, 2] = 2, -1  # dst(x,y) <- src(x+2, y-1)
    out = O.warp_frame(img, M)
    assert np.allclose(out[1:, :-2], img[:-1, 2:])
    assert np.all(out[0] == 0) and np.all(out[:, -2:] == 0)  # BORDER_CONSTANT 0


def test_warp_matches_cv2_in_quantised_mode():
    """cv2.warpPerspective == bilinear with source coordinates rounded to 1/32 px (Appendix A-20)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.random((40, 56, 3)).astype(np.float32)
    M = np.eye(3)
    M[:2, :2] += rng.uniform(-0.01, 0.01, (2, 2))
    M[:2, 2] += rng.uniform(-4, 4, 2)
    M[2, :2] += rng.uniform(-1e-5, 1e-5, 2)
    ref = cv2.warpPerspective(img, M, (56, 40), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
    got = O.warp_frame(img, M, quantize_1_32=True)
    assert np.abs(got - ref).max() < 2e-6
    assert np.abs(O.warp_frame(img, M) - ref).max() > 1e-4  # exact mode really differs from cv2's quantisation


def test_tiling_divide_merge_roundtrip():
    x = torch.rand(1, 3, 2, 50, 70)
    tiles = O.tensor_divide_burst(x, 20, 10)
    assert tiles.shape == (3 * 4, 3, 2, 40, 40)
    # centre of every tile is the (reflect padded) source; stitching the centres of frame 0 gives it back
    back = O.tensor_merge(tiles[:, 0], (50, 70), psize=20, overlap=10)
    assert torch.equal(back[0], x[0, 0])
    # halo is a reflection: first tile's top-left corner mirrors the image
    assert torch.equal(tiles[0, 0, 0, 0, 10:30], x[0, 0, 0, 10, 0:20])


def test_faf_gate_algebraic_identity():
    """DESIGN.md 'FAF gate identity': |sum_c(E_f - R) - sum_c(E_0 - R)| == |wsum * (x_f - x_0)| (float64)."""
    fu = O.FAFBlock(8, 4).double()
    feat = torch.rand(2, 4, 8, 9, 11, dtype=torch.float64)
    _, gate = fu.guided(feat)
    wsum = fu.temporal_attn1.weight.sum(0, keepdim=True)
    s = F.conv2d((feat[:, 1:] - feat[:, :1]).reshape(6, 8, 9, 11), wsum, padding=1).view(2, 3, 9, 11)
    assert torch.allclose(gate, torch.sigmoid(s.abs()), atol=1e-12)


def test_forward_shapes_batch_independence_and_determinism():
    cfg = dict(num_frames=3, img_size=20, in_channels=3, embed_dim=32, window_length=10)
    m = O.build_oracle(0, **cfg)
    x = torch.rand(2, 3, 3, 20, 20, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = m(x)
        y0 = m(x[:1])
    assert y.shape == (2, 3, 80, 80)
    assert torch.allclose(y[:1], y0, atol=1e-6)
    assert torch.equal(O.build_oracle(0, **cfg).head.weight, m.head.weight)
    with pytest.raises(AssertionError):
        m(torch.rand(1, 3, 3, 24, 24))


def test_golden_small_model():
    """Committed fixture (tests/golden/make_golden.py): guards the oracle itself against drift."""
    g = torch.load(os.path.join(GOLD, "small_model.pt"))
    m = O.build_oracle(g["seed"], **g["cfg"])
    with torch.no_grad():
        st = m.forward_stages(g["x"])
    assert torch.allclose(st["out"], g["out"], atol=2e-5)
    for k, (mean, std) in g["stage_stats"].items():
        assert abs(st[k].mean().item() - mean) < 1e-4 and abs(st[k].std().item() - std) < 1e-4, k


def test_golden_warp():
    g = np.load(os.path.join(GOLD, "warp.npz"))
    out = O.warp_burst(g["burst"], g["M"])
    assert np.abs(out - g["out"]).max() < 1e-12


def test_flow_registration_oracle_matches_scipy_map_coordinates():
    """8f-4 pin: the restated jax map_coordinates(order=1, mode="nearest") of registration/optical_flow/register.py:11-47 against
    scipy.ndimage.map_coordinates evaluated in float64 (the API jax.scipy mirrors).  Tolerance = fp32 rounding of the
    coordinates (the reference computes them in fp32)."""
    import numpy as np
    import scipy.ndimage as ndi
    from oracle.fbanet_oracle import flow_register_burst, flow_register_frame
    rng = np.random.default_rng(0)
    for (H, W, C, amp) in [(37, 53, 3, 12.0), (80, 80, 4, 3.0), (9, 7, 1, 30.0)]:
        f = rng.random((H, W, C)).astype(np.float32)
        fl = ((rng.random((H, W, 2)) - 0.5) * amp).astype(np.float32)
        gy, gx = np.mgrid[:H, :W]
        ref = np.stack([ndi.map_coordinates(f[..., c].astype(np.float64), [gy - fl[..., 0].astype(np.float64), gx - fl[..., 1].astype(np.float64)],
                                            order=1, mode="nearest") for c in range(C)], -1)
        assert np.abs(flow_register_frame(f, fl) - ref).max() < 5e-6
    # integer flows are exact shifts with edge replication; the reference frame passes through
    f = rng.random((3, 8, 10, 2)).astype(np.float32)
    fl = np.zeros((2, 8, 10, 2), np.float32)
    fl[..., 0], fl[..., 1] = 1.0, -2.0
    out = flow_register_burst(f, fl)
    assert np.array_equal(out[0], f[0])
    yy, xx = np.clip(np.arange(8) - 1, 0, 7), np.clip(np.arange(10) + 2, 0, 9)
    assert np.array_equal(out[1], f[1][yy][:, xx])


def _ecc_pair(seed, H=96, W=128, noise=0.01):
    import cv2
    import numpy as np
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.random((H, W)).astype(np.float32), (0, 0), 2.5)
    img = (img - img.min()) / (img.max() - img.min())
    M = np.eye(3, dtype=np.float32)
    M[0, 2], M[1, 2] = rng.uniform(-3, 3), rng.uniform(-3, 3)
    M[0, 1], M[1, 0] = rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01)
    M[2, 0] = rng.uniform(-1e-5, 1e-5)
    inp = cv2.warpPerspective(img, M, (W, H), flags=cv2.INTER_LINEAR) + rng.normal(0, noise, (H, W)).astype(np.float32)
    return img, inp.astype(np.float32)


def test_ecc_oracle_matches_cv2():
    """8f-4 pin: the restated forward-additive ECC (homography_alignment.py:19-45 -> cv2.findTransformECC, MOTION_HOMOGRAPHY, 100
    iterations, eps 1e-10, gaussFiltSize 5) against the cv2 in this container, stage by stage and end to end.  The end-to-end
    tolerance (0.01 px over the image, 5e-4 in rho) is OpenCV's own 1/32-px coordinate quantisation in warpPerspective."""
    import cv2
    import numpy as np
    from oracle.fbanet_oracle import bgr2gray, ecc_blur5, ecc_gradients, ecc_homography, homography_coord_diff as _coord_diff
    rng = np.random.default_rng(0)
    a = rng.random((45, 61)).astype(np.float32)
    assert np.abs(ecc_blur5(a) - cv2.GaussianBlur(a, (5, 5), 0)).max() < 5e-7
    gx, gy = ecc_gradients(a)
    assert np.abs(gx - cv2.filter2D(a, -1, np.array([[-0.5, 0, 0.5]], np.float32))).max() < 1e-7
    assert np.abs(gy - cv2.filter2D(a, -1, np.array([[-0.5], [0], [0.5]], np.float32))).max() < 1e-7
    c = rng.random((20, 30, 3)).astype(np.float32)
    assert np.abs(bgr2gray(c) - cv2.cvtColor(c, cv2.COLOR_BGR2GRAY)).max() < 5e-7
    crit = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-10)
    for seed in range(3):
        t, i = _ecc_pair(seed)
        cc, wm = cv2.findTransformECC(t, i, np.eye(3, dtype=np.float32), cv2.MOTION_HOMOGRAPHY, crit)
        rho, M = ecc_homography(t, i)
        assert abs(cc - rho) < 5e-4, (cc, rho)
        assert _coord_diff(wm, M, *t.shape) < 1e-2


def test_tiling_oracle_matches_vectors_from_the_reference_code():
    """8f-1 pin: `tests/golden/tiling_reference.npz` was produced by executing the reference's own
    `utils/dataset_utils.py::tensor_divide_burst / tensor_merge` (tests/golden/make_golden_reference.py); the restatement must
    reproduce it bit for bit."""
    import numpy as np
    from oracle.fbanet_oracle import tensor_divide_burst, tensor_merge
    from tests_golden_helpers import tiling_reference
    d, sr = tiling_reference()
    ps, ov, sc = int(d["psize"]), int(d["overlap"]), int(d["scale"])
    burst = torch.from_numpy(d["burst"])
    assert np.array_equal(tensor_divide_burst(burst, ps, ov).numpy(), d["tiles"])
    H, W = burst.shape[-2:]
    assert np.array_equal(tensor_merge(torch.from_numpy(sr), (sc * H, sc * W), sc * ps, sc * ov).numpy(), d["merged"])


def test_training_loss_oracle_matches_vectors_from_the_reference_code():
    """8f-3 pin: value and autograd gradient of the restated CharbonnierLoss + 3 * GWLoss against `tests/golden/loss_reference.npz`,
    produced by executing the reference's own `losses.py` (tests/golden/make_golden_reference.py)."""
    import os
    import numpy as np
    from oracle.fbanet_oracle import charbonnier_loss, gw_loss, training_loss
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_reference.npz"))
    x, y = torch.from_numpy(d["x"]).requires_grad_(True), torch.from_numpy(d["y"])
    assert abs(charbonnier_loss(x, y).item() - float(d["charbonnier"])) < 1e-7
    assert abs(gw_loss(x, y).item() - float(d["gw"])) < 1e-6
    total = training_loss(x, y)
    total.backward()
    assert abs(total.item() - float(d["total"])) < 1e-6
    assert np.abs(x.grad.numpy() - d["grad"]).max() < 1e-7
    # the trainer's two lines as written (train.py.bak:167-168): clamp(restored, 0, 1) in front of both criteria
    xc = torch.from_numpy(d["x"]).requires_grad_(True)
    tc = training_loss(xc, y, clamp_restored=True)
    tc.backward()
    assert abs(tc.item() - float(d["total_clamped"])) < 1e-6
    assert np.abs(xc.grad.numpy() - d["grad_clamped"]).max() < 1e-7
    assert np.abs(d["grad_clamped"] - d["grad"]).max() > 1e-6                # the fixture does leave [0, 1]: the two differ
