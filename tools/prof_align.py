#!/usr/bin/env python
"""SURVEY 8f-4 measurement: the on-GPU alignment front end (ECC homography estimation + K1 warp) on the cfg2 burst shape, with
`cv2.findTransformECC` + `cv2.warpPerspective` (what `homography_alignment.py:19-55` runs per frame) timed beside it on the host
cores over a bounded sample of the same pairs, and the agreement of the two sets of matrices.  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbanet_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bursts", type=int, default=64)
    ap.add_argument("--size", type=int, default=160)
    ap.add_argument("--cpu-pairs", type=int, default=26)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T, S = a.bursts, 14, a.size
    g = torch.Generator().manual_seed(0)
    # smooth random base frames (noise blurred by repeated 3x3 box filters), frames = known homographies of the base + noise
    base = torch.rand(B, 1, 3, S + 16, S + 16, generator=g)
    k = torch.ones(3, 1, 5, 5) / 25.0
    for _ in range(3):
        base = torch.nn.functional.conv2d(torch.nn.functional.pad(base[:, 0], (2, 2, 2, 2), mode="reflect"), k, groups=3)[:, None]
    base = base[..., 8:-8, 8:-8].contiguous()
    base = (base - base.amin()) / (base.amax() - base.amin())
    Mt = torch.eye(3, dtype=torch.float64).repeat(B, T, 1, 1)
    Mt[:, 1:, :2, 2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 6 - 3
    Mt[:, 1:, 0, 1] = torch.rand(B, T - 1, generator=g, dtype=torch.float64) * 0.02 - 0.01
    Mt[:, 1:, 1, 0] = torch.rand(B, T - 1, generator=g, dtype=torch.float64) * 0.02 - 0.01
    Mt[:, 1:, 2, 0] = torch.rand(B, T - 1, generator=g, dtype=torch.float64) * 2e-5 - 1e-5
    burst = ops.warp_burst(base.expand(B, T, 3, S, S).contiguous().to(dev), Mt)
    burst = (burst + 0.01 * torch.randn(burst.shape, generator=torch.Generator(device=dev).manual_seed(1), device=dev)).contiguous()
    burst[:, 0] = base[:, 0].to(dev)

    def gpu_once():
        M, rho, it = ops.ecc_homography_burst(burst)
        out = ops.warp_burst(burst, M)
        return M, rho, it, out

    gpu_once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        M, rho, it, out = gpu_once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    pairs = B * (T - 1)
    # CPU: cv2 on a bounded sample of the same pairs (whole bursts first)
    import cv2
    hb = burst.permute(0, 1, 3, 4, 2).contiguous().cpu().numpy()
    crit = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-10)
    n_cpu, worst, t0 = 0, 0.0, time.perf_counter()
    Y, X = np.mgrid[:S, :S].astype(np.float64)

    def co(Mx):
        d = Mx[2, 0] * X + Mx[2, 1] * Y + Mx[2, 2]
        return (Mx[0, 0] * X + Mx[0, 1] * Y + Mx[0, 2]) / d, (Mx[1, 0] * X + Mx[1, 1] * Y + Mx[1, 2]) / d
    Mh = M.cpu().numpy()
    cpu_ecc_s = 0.0
    for b in range(B):
        g0 = cv2.cvtColor(hb[b, 0], cv2.COLOR_BGR2GRAY)
        for t in range(1, T):
            if n_cpu >= a.cpu_pairs:
                break
            t1 = time.perf_counter()
            _, wm = cv2.findTransformECC(g0, cv2.cvtColor(hb[b, t], cv2.COLOR_BGR2GRAY), np.eye(3, dtype=np.float32), cv2.MOTION_HOMOGRAPHY, crit)
            cpu_ecc_s += time.perf_counter() - t1
            cv2.warpPerspective(hb[b, t], wm, (S, S), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
            (p, q), (r, s) = co(wm.astype(np.float64)), co(Mh[b, t])
            worst = max(worst, np.abs(p - r).max(), np.abs(q - s).max())
            n_cpu += 1
    cpu_s = time.perf_counter() - t0
    res = {"workload": f"{B} bursts x {T - 1} frame pairs, {S}x{S} RGB: ECC homography (100 iterations max, eps 1e-10) + bilinear warp",
           "gpu_ms": ms, "gpu_pairs_per_s": pairs / (ms / 1e3), "gpu_iterations_mean": float(it[:, 1:].float().mean().item()),
           "gpu_failed_pairs": int((it[:, 1:] < 0).sum().item()), "gpu_rho_mean": float(rho[:, 1:].mean().item()),
           "cpu_kind": "reference (cv2.findTransformECC + cv2.warpPerspective, cv2 " + cv2.__version__ + ")", "cpu_cores": os.cpu_count(),
           "cpu_sample_pairs": n_cpu, "cpu_pairs_per_s": n_cpu / cpu_s, "cpu_ecc_ms_per_pair": 1e3 * cpu_ecc_s / max(1, n_cpu),
           "max_coord_diff_px_gpu_vs_cv2_on_sample": float(worst)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
