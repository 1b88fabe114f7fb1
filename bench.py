#!/usr/bin/env python
"""Headline benchmark: FBANet BaseModel burst-SR forward, bursts/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU forward (oracle port)

One "step" = one forward over one batch of synthetic bursts (cfg2: batch 64 x 14x160x160 RGB per GPU,
embed_dim 64, win 10, random-init weights).  N>1: one process per GPU under torchrun, bursts sharded
across ranks, NO data-path collective (weak scaling); timing = max over ranks of CUDA-event time.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10)
MP_PER_BURST = (4 * CFG["img_size"]) ** 2 / 1e6  # 0.4096 output megapixels per burst
WORKLOAD = "cfg2: BaseModel embed_dim=64 win=10 inference, synthetic 14x160x160 RGB bursts -> 640x640 (x4)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.samples.append([s.strip() for s in r.stdout.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def cpu_forward_rate(threads: int, reps: int, seed: int = 0):
    """Reference CPU forward (oracle port) on the host cores: bursts/s over `reps` single-burst forwards."""
    from oracle.fbanet_oracle import build_oracle  # the one place bench.py may execute oracle/
    torch.set_num_threads(threads)
    o = build_oracle(seed, **CFG)
    x = torch.rand(1, CFG["num_frames"], CFG["in_channels"], CFG["img_size"], CFG["img_size"], generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        o(x)  # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            o(x)
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1.0 / ts[len(ts) // 2], ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle.fbanet_oracle import build_oracle
    o = build_oracle(0, **CFG)
    x = torch.rand(1, CFG["num_frames"], CFG["in_channels"], CFG["img_size"], CFG["img_size"], generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            o(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o(x)
        dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} steps x 1 burst (14x160x160 RGB) of the batch-64 workload, torch CPU fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "bursts_per_sec", "value": v, "unit": "bursts/s", "output_mp_per_s": v * MP_PER_BURST,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": 1, "note": "reference JAX forward is not executable (SURVEY F2/F3); CPU oracle port timed"},
        "cpu_baseline": {"value": v, "unit": "bursts/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "bursts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] beside the headline: computed AFTER the cfg2 timed regions (the headline is unchanged by them), each
# device-timed with a barrier + synchronize on both sides and the max over ranks, reported under "other_configs".
# ------------------------------------------------------------------------------------------------------------------------------
def _max_ms(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def other_cfg3(dev, stream, rank, world, steps, dtype):
    """configs[2]: RealBSR-RAW shape -- batch 64 of 14x4x80x80 packed-Bayer bursts per GPU, homography warp (K1, non-zero
    perspective rows) + FAF fusion + SR x4, burst-sharded, CUDA-graph replay."""
    import torch.distributed as dist
    from fbanet_b200 import BaseModel, ops
    cfg = dict(num_frames=14, img_size=80, in_channels=4, embed_dim=64, window_length=10)
    B, T = 64, 14
    model = BaseModel(**cfg, token_projection="linear", token_mlp="leff", dtype=dtype, seed=0).to(dev).eval()
    x = torch.rand(B, T, 4, 80, 80, generator=torch.Generator().manual_seed(2000 + rank)).to(dev)
    g = torch.Generator().manual_seed(1)
    M = torch.eye(3, dtype=torch.float64).repeat(B, T, 1, 1)
    M[:, 1:, :2, 2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 8 - 4
    M[:, 1:, :2, :2] += torch.rand(B, T - 1, 2, 2, generator=g, dtype=torch.float64) * 0.02 - 0.01
    M[:, 1:, 2, :2] = torch.rand(B, T - 1, 2, generator=g, dtype=torch.float64) * 2e-5 - 1e-5
    Md = M.to(dev)
    with torch.cuda.stream(stream):
        step = lambda: model(x, homographies=Md)   # K1 fused into the head conv's sampling on the tensor-core path (no warp launch)
        for _ in range(3):
            step()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            step()
        graph.replay()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            graph.replay()
        e1.record(stream)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        ms = _max_ms(e0.elapsed_time(e1) / steps, dev, world)
        fam = ops.profile_ops(step, stream, by_tag=False)
    warp_ms = fam.get("fbanet_warp_sm100", (0.0, 0, 0))[0]   # 0 when the warp is fused into the head conv
    head_ms = fam.get("fbanet_head_conv_sm100", (0.0, 0, 0))[0]
    return {"workload": "cfg3: RealBSR-RAW shape, 14x4x80x80 packed-Bayer bursts, per-frame homography warp + FAF fusion + SR x4 -> 320x320",
            "batch_per_gpu": B, "n_gpus": world, "dtype": dtype, "steps": steps, "ms_per_step": ms, "bursts_per_s": B * world / (ms / 1e3),
            "output_mp_per_s": B * world * 0.1024 / (ms / 1e3), "cuda_graph": True, "warp_fused_into_head_conv": warp_ms == 0.0, "head_conv_ms": head_ms,
            "warp_ms": warp_ms,
            "warp_hbm_gbs": (2 * x.numel() * 4 / 1e9) / (warp_ms / 1e3) if warp_ms else None}


def other_cfg4(dev, rank, world, steps, dtype):
    """configs[3]: one full-size 14x3x1080x1920 burst x4 -> 4320x7680, row-band sharded over the ranks; the 40-px tile halo is read
    from the neighbours' bands over NVLink inside the tile-gather kernel (symmetric memory, no NCCL on the data path); the stitched
    image is compared bit for bit with the single-GPU replicated-burst driver on rank 0."""
    import torch.distributed as dist
    from fbanet_b200 import BaseModel
    from fbanet_b200.dist import band_rows, halo_sources, shard_range
    from fbanet_b200.tiling import BandedSession, infer_full_resolution, infer_full_resolution_banded
    H, W, T, C = 1080, 1920, 14, 3
    model = BaseModel(num_frames=T, img_size=160, in_channels=C, embed_dim=64, window_length=10, token_projection="linear",
                      token_mlp="leff", dtype=dtype, seed=0).to(dev).eval()
    row0 = band_rows(H, world)
    full = torch.rand(T, C, H, W, generator=torch.Generator().manual_seed(0))      # every rank draws the same image, uploads ITS band
    band = full[:, :, row0[rank]:row0[rank + 1]].contiguous().to(dev)
    session = BandedSession(T, C, H, W, dev)
    step = lambda: infer_full_resolution_banded(model, band, H, W, tile_batch=64, gather_to=0, session=session)
    out = step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize(dev)
    dist.barrier()
    ms = _max_ms(e0.elapsed_time(e1) / steps, dev, world)
    nh, nw = -(-H // 80), -(-W // 80)
    t0, t1 = shard_range(nh * nw, rank, world)
    need = halo_sources(H, 80, 40, (t0 // nw, (t1 - 1) // nw + 1), row0)
    remote = torch.tensor([sum(v for k, v in need.items() if k != rank) * W * T * C * 4], dtype=torch.float64, device=dev)
    dist.all_reduce(remote)
    res = {"workload": f"cfg4: {T}x{C}x{H}x{W} burst x4 -> {4 * H}x{4 * W}, {nh * nw} tiles of 160x160 (psize 80, overlap 40), row-band sharded",
           "n_gpus": world, "dtype": dtype, "steps": steps, "ms_per_image": ms, "output_mp_per_s": 16 * H * W / 1e6 / (ms / 1e3),
           "halo_bytes_read_from_peers_per_image": int(remote.item()),
           "data_path": "peer loads/stores on symmetric memory inside the tile gather / stitch kernels; no NCCL collective"}
    if rank == 0:
        ref = infer_full_resolution(model, full[None].to(dev), tile_batch=64)
        torch.cuda.synchronize(dev)
        res["bit_identical_to_single_gpu_driver"] = bool(torch.equal(out, ref))
        del ref
    dist.barrier()
    return res


def other_cfg5(dev, rank, world, steps, dtype):
    """configs[4]: one data-parallel training step -- 2 bursts per GPU, 14x3x160x160 -> 640x640 target, clamp + Charbonnier + 3 GW
    (train.py.bak:167-168), backward on the C-ABI bricks with the bucketed NCCL gradient all-reduce overlapping it, fused AdamW."""
    import torch.distributed as dist
    from fbanet_b200 import BaseModel, ops, train
    model = BaseModel(**CFG, token_projection="linear", token_mlp="leff", dtype=dtype, seed=0).to(dev)
    for p in model.parameters():
        p.requires_grad_(True)
    flat = train.FlatParams(model.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    burst = torch.rand(2, 14, 3, 160, 160, generator=g).to(dev)
    target = torch.rand(2, 3, 640, 640, generator=g).to(dev)
    gen = torch.Generator().manual_seed(7 + rank)
    first = train.train_step(model, flat, burst, target, 1e-4, generator=gen)[0].item()
    train.train_step(model, flat, burst, target, 1e-4, generator=gen)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    flat.reduce_events = []
    before = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = train.train_step(model, flat, burst, target, 1e-4, generator=gen)
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = _max_ms(e0.elapsed_time(e1) / steps, dev, world)
    exposed = sum(a.elapsed_time(b) for a, b in flat.reduce_events) / max(1, len(flat.reduce_events))
    flat.reduce_events = None
    return {"workload": "cfg5: training step, 2 bursts/GPU 14x3x160x160 -> 640x640, clamp + Charbonnier + 3 GW, AdamW(1e-4, wd 0.02)",
            "n_gpus": world, "global_batch": 2 * world, "dtype": dtype, "steps": steps, "ms_per_step": ms, "bursts_per_s": 2 * world / (ms / 1e3),
            "collective": f"NCCL all-reduce(sum) of {flat.numel * 4 / 1e6:.1f} MB fp32 gradients in {len(flat._buckets)} buckets, issued during the backward" if world > 1 else "none (1 rank)",
            "allreduce_exposed_ms": _max_ms(exposed, dev, world), "loss_first": first, "loss_last": loss[0].item(),
            "launches_per_step": (ops.LAUNCHES - before) // steps, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}


def run_other_configs(args, dev, stream, rank, world):
    import gc
    out = {}
    jobs = [("cfg3", lambda: other_cfg3(dev, stream, rank, world, max(3, min(args.steps, 10)), args.dtype))]
    if world > 1 or args.other_configs == "all":
        if world > 1:
            jobs.append(("cfg4", lambda: other_cfg4(dev, rank, world, 2, args.dtype)))
        jobs.append(("cfg5", lambda: other_cfg5(dev, rank, world, 3, args.dtype)))
    for name, fn in jobs:
        try:
            out[name] = fn()
        except Exception as e:   # these lines are extras: a failure is reported, never fatal to the headline
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        gc.collect()
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="bursts per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-impl", default="auto", choices=["auto", "simt"])
    ap.add_argument("--breakdown", action="store_true", help="print a per-kernel CUDA-event breakdown of one step to stderr")
    ap.add_argument("--host-chunk", type=int, default=0, help="bursts per pipelined chunk of the end-to-end (host buffer) call; 0 = model default")
    ap.add_argument("--other-configs", default="auto", choices=["auto", "none", "all"],
                    help="configs[2..4] beside the headline: auto = cfg3 always, cfg4 + cfg5 when WORLD_SIZE > 1; all = cfg5 on one GPU too")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from fbanet_b200 import BaseModel, ops, _lib as L

    # stdout carries exactly ONE JSON line: library banners (NCCL prints its version to stdout) go to stderr until then
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    B = args.batch

    model = BaseModel(**CFG, token_projection="linear", token_mlp="leff", dtype=args.dtype, seed=0,
                      impl=L.IMPL_SIMT if args.kernel_impl == "simt" else L.IMPL_AUTO).to(dev).eval()
    # burst sharding: rank r owns global bursts [r*B, (r+1)*B)  (pipeline/real_bsr_dataset.py:82-83 rule)
    g = torch.Generator().manual_seed(1000 + rank)
    host_in = torch.rand(B, CFG["num_frames"], CFG["in_channels"], CFG["img_size"], CFG["img_size"], generator=g).pin_memory()
    host_out = torch.empty(B, CFG["in_channels"], 4 * CFG["img_size"], 4 * CFG["img_size"]).pin_memory()
    x = host_in.to(dev)
    stream = torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        # ---- warm-up (also packs weights), kernel-level roofline instrumentation on the last warm-up ----
        for _ in range(W):
            y = model(x)
        stream.synchronize()
        ops.LAUNCHES = 0
        model(x)
        launches_per_step = ops.LAUNCHES
        stream.synchronize()

        # ---- device-resident timed region ----
        graph = None
        if not args.no_graph:
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    y = model(x)
                graph.replay()
                stream.synchronize()
            except Exception as e:  # capture is an optimisation, never a correctness dependency
                print(f"[bench] CUDA graph capture failed ({e}); running eagerly", file=sys.stderr)
                graph = None
        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                y = model(x)
        e1.record(stream)
        stream.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)

        # ---- dominant-kernel roofline: CUDA events around every conv/GEMM launch of one more eager pass ----
        prof = ops.profile_conv_gemm(lambda: model(x), stream)

        # ---- per-kernel CUDA-event pass: the bandwidth-bound kernels against the HBM roofline (algorithmic bytes / time) ----
        fam = ops.profile_ops(lambda: model(x), stream, by_tag=False)
        # K1 homography warp: not part of the cfg2 step (its bursts are pre-aligned), measured on the same burst shape
        Mh = torch.eye(3, dtype=torch.float64).repeat(B, CFG["num_frames"], 1, 1)
        Mh[:, 1:, :2, 2] = torch.rand(B, CFG["num_frames"] - 1, 2, generator=torch.Generator().manual_seed(1), dtype=torch.float64) * 8 - 4
        Mh = Mh.to(dev)
        ops.warp_burst(x, Mh)
        fam.update(ops.profile_ops(lambda: [ops.warp_burst(x, Mh) for _ in range(3)], stream, by_tag=False))
        # 8f-4 optical-flow registration (registration/optical_flow/register.py), same burst shape, +-4 px synthetic flow
        flow = (torch.rand(B, CFG["num_frames"] - 1, CFG["img_size"], CFG["img_size"], 2, generator=torch.Generator().manual_seed(2)) * 8 - 4).to(dev)
        ops.flow_warp_burst(x, flow)
        fam.update(ops.profile_ops(lambda: [ops.flow_warp_burst(x, flow) for _ in range(3)], stream, by_tag=False))
        if args.breakdown and rank == 0:
            bd = ops.profile_ops(lambda: model(x), stream)
            tot = sum(v[0] for v in bd.values())
            print(f"[breakdown] one step, eager, CUDA events: {tot:.2f} ms", file=sys.stderr)
            for k, (t_ms, n, by) in sorted(bd.items(), key=lambda kv: -kv[1][0]):
                print(f"[breakdown] {t_ms:9.3f} ms {100 * t_ms / tot:5.1f}%  n={n:3d}  {by / t_ms / 1e9 if t_ms else 0:6.2f} TB/s alg  {k}", file=sys.stderr)

        # ---- end to end through the public API with HOST buffers (pinned H2D + D2H inside the timed region) ----
        # headline: the streaming call a loop over batches makes (`infer_host_stream`: every step uploads its own 64 bursts from
        # pinned host memory and downloads its own SR images into one of two pinned host buffers; the copies of neighbouring
        # steps overlap the forward).  `e2e_blocking`: one blocking `infer_host` call per step (first upload / last download exposed).
        if args.host_chunk:
            model.host_chunk = args.host_chunk
        host_outs = [host_out, torch.empty_like(host_out).pin_memory()]

        def stream_steps(n, src, dsts, out_dtype):
            k = 0
            for _ in model.infer_host_stream((src for _ in range(n)), outs=(dsts[i & 1] for i in range(n)), out_dtype=out_dtype, depth=2):
                k += 1
            assert k == n

        stream_steps(max(3, args.warmup), host_in, host_outs, torch.float32)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        stream_steps(args.steps, host_in, host_outs, torch.float32)   # returns after the last download has landed
        h1.record(stream)
        stream.synchronize()
        barrier()
        ms_e2e = h0.elapsed_time(h1)

        for _ in range(2):
            model.infer_host(host_in, host_out)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for _ in range(args.steps):
            model.infer_host(host_in, host_out)
        b1.record(stream)
        stream.synchronize()
        barrier()
        ms_e2e_blocking = b0.elapsed_time(b1)

        # ---- the same streaming call with the reference's own 8-bit data path: uint8 frames in (normalised on the device,
        # train.py:82-83), uint8 SR image out (clamp + ToPILImage truncation on the device, test_in_any_resolution.py:93-101) ----
        host_in_u8 = (host_in * 255.0).to(torch.uint8).pin_memory()
        host_outs_u8 = [torch.empty(host_out.shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        stream_steps(3, host_in_u8, host_outs_u8, torch.uint8)
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        stream_steps(args.steps, host_in_u8, host_outs_u8, torch.uint8)
        n1.record(stream)
        stream.synchronize()
        barrier()
        ms_e2e_u8 = n0.elapsed_time(n1)

    t = torch.tensor([ms, ms_e2e, ms_e2e_u8, ms_e2e_blocking], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_u8, ms_e2e_blocking = t.tolist()
    others, used_graph = None, graph is not None
    if args.other_configs != "none" and args.dtype == "bf16":
        used_graph = graph is not None
        del graph, y, x
        model._host_graphs.clear()
        torch.cuda.empty_cache()
        others = run_other_configs(args, dev, stream, rank, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sys.stdout.flush()
    os.dup2(real_stdout, 1)

    pk, pk_src = peaks()
    total_bursts = B * world * args.steps
    value = total_bursts / (ms / 1e3)
    e2e = total_bursts / (ms_e2e / 1e3)
    peak_tf = pk["bf16_tflops_sustained"]
    # DRAM traffic of the dominant kernel family from the committed ncu launch list of this same command
    # (profiles/*_kernel_summary_*.json, tools/launch_summary.py): bytes per launch, like `achieved`
    traffic, traffic_src = None, None
    summ = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if "kernel_summary" in f and f.endswith(".json")) if os.path.isdir(os.path.join(ROOT, "profiles")) else []
    if summ:
        sj = json.load(open(os.path.join(ROOT, "profiles", summ[-1])))
        cf = sj.get("families", {}).get("conv_gemm_tcgen05_kernel")
        order, per_launch = prof.get("order", []), sj.get("conv_gemm_launches", [])
        if per_launch and len(per_launch) == len(order):
            # the ncu launch list and this run launch the same sequence of implicit GEMMs: take the 3x3 / 4x4 convolutions' share
            sel = [d["dram_bytes"] for d, taps in zip(per_launch, order) if taps > 1]
            traffic = sum(sel) / max(1, len(sel))
            traffic_src = (f"profiles/{summ[-1]} (ncu dram__bytes_read+write per launch, the {len(sel)} conv launches of one batch-"
                           f"{sj.get('batch', 64)} step, matched by launch order)")
        elif cf and cf.get("launches"):
            traffic = cf["dram_bytes"] / cf["launches"]
            traffic_src = f"profiles/{summ[-1]} (ncu dram__bytes_read+write, all {cf['launches']} implicit-GEMM launches of one batch-{sj.get('batch', 64)} step)"
    step_ms = ms / args.steps
    cls = prof.get("classes", {})
    conv = cls.get("conv", {"ms": 0.0, "flops": 0.0, "bytes": 0, "launches": 0})
    g1 = cls.get("gemm1x1", {"ms": 0.0, "flops": 0.0, "bytes": 0, "launches": 0})
    conv_tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    hbm_peak = pk["hbm_gbs"]
    # the dominant kernel: the tcgen05 implicit GEMM.  Its 3x3 / 4x4 convolutions are tensor-pipe work and are judged against the
    # bf16 peak; its 1x1 GEMMs (qkv / proj / fc1 / fusion / transposed convs) move ~100 FLOP per byte at these channel counts and
    # are judged against HBM (roofline_classes, gemm1x1_shapes) -- one averaged "tensor" fraction hid which ones were bad.
    roof = {
        "bound": "tensor", "kernel": "fbanet_conv_gemm_sm100, 3x3 / 4x4 convolutions (implicit GEMM on tcgen05, all such launches of one step)",
        "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": conv["bytes"] / max(1, conv["launches"]),
        "peak_source": pk_src + ", sustained bf16 (kernel timed inside a long step)",
        "launches": conv["launches"], "ms": conv["ms"], "share_of_step": conv["ms"] / step_ms, "flops_per_step": conv["flops"],
        "whole_family": {"launches": prof["launches"], "ms": prof["ms"], "tflops": prof["tflops"], "frac_of_tensor_peak": prof["tflops"] / peak_tf,
                         "share_of_step": prof["ms"] / step_ms, "flops_per_step": prof["flops"]},
    }
    g1_gbs = g1["bytes"] / g1["ms"] / 1e6 if g1["ms"] > 0 else 0.0
    roofline_classes = [
        {"class": "conv3x3/4x4 (implicit GEMM)", "bound": "tensor", "launches": conv["launches"], "ms": conv["ms"], "achieved": conv_tf, "peak": peak_tf,
         "unit": "TFLOP/s", "frac": conv_tf / peak_tf},
        {"class": "1x1 GEMMs (linear / 1x1 / transposed conv)", "bound": "hbm", "launches": g1["launches"], "ms": g1["ms"], "achieved": g1_gbs,
         "peak": hbm_peak, "unit": "GB/s", "frac": g1_gbs / hbm_peak, "tflops": g1["flops"] / (g1["ms"] * 1e-3) / 1e12 if g1["ms"] > 0 else 0.0},
    ]
    shapes = prof.get("shapes", {})
    # 1x1 GEMMs that compute their input's LayerNorm inside (shape tag "... ln"): the LayerNorm's time joins the class while its own pass and
    # bytes leave the step, so the class above is not comparable with earlier builds -- split it
    ln_sh = [v for t, v in shapes.items() if v["taps"] == 1 and v["ms"] > 0 and t.endswith(" ln")]
    pl_sh = [v for t, v in shapes.items() if v["taps"] == 1 and v["ms"] > 0 and not t.endswith(" ln")]
    for label, grp in (("1x1 GEMMs without LayerNorm inside", pl_sh), ("1x1 GEMMs with their LayerNorm inside (LayerNorm pass saved: rows x C x 4 bytes each)", ln_sh)):
        if grp:
            g_ms, g_by = sum(v["ms"] for v in grp), sum(v["bytes"] for v in grp)
            roofline_classes.append({"class": label, "bound": "hbm", "launches": sum(v["launches"] for v in grp), "ms": g_ms, "achieved": g_by / g_ms / 1e6,
                                     "peak": hbm_peak, "unit": "GB/s", "frac": g_by / g_ms / 1e6 / hbm_peak})
    gemm1x1_shapes = [{"shape": t, "launches": v["launches"], "ms": v["ms"], "achieved": v["bytes"] / v["ms"] / 1e6, "unit": "GB/s",
                       "frac": v["bytes"] / v["ms"] / 1e6 / hbm_peak} for t, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]) if v["taps"] == 1 and v["ms"] > 0]
    conv_shapes = [{"shape": t, "launches": v["launches"], "ms": v["ms"], "achieved": v["flops"] / (v["ms"] * 1e-3) / 1e12, "unit": "TFLOP/s",
                    "frac": v["flops"] / (v["ms"] * 1e-3) / 1e12 / peak_tf} for t, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]) if v["taps"] > 1 and v["ms"] > 0]
    hbm_kernels = []
    for name in ("fbanet_warp_sm100", "fbanet_flow_warp_sm100", "fbanet_head_conv_sm100", "fbanet_layernorm_sm100", "fbanet_assemble_sm100"):
        if name in fam and fam[name][0] > 0:
            t_ms, n, by = fam[name]
            gbs = by / t_ms / 1e6
            hbm_kernels.append({"kernel": name, "bound": "hbm", "launches": n, "ms": t_ms, "algorithmic_bytes": by, "achieved": gbs, "peak": hbm_peak,
                                "unit": "GB/s", "frac": gbs / hbm_peak})
    # K2 as SURVEY 8(d) defines it: the gate (score conv 64 -> 1 + sigmoid gate) AND the K = 896 1x1 fusion, (2F+2) H W E s bytes per burst
    Fr, S, E = CFG["num_frames"], CFG["img_size"], CFG["embed_dim"]
    k2_parts = {"score_conv": sum(v["ms"] for t, v in shapes.items() if t.startswith("k3s1 64->16 @%dx%d " % (S, S)) and t.endswith("st4")),
                "gate_apply": fam.get("fbanet_faf_gate_sm100", (0.0, 0, 0))[0],
                "fuse_1x1": sum(v["ms"] for t, v in shapes.items() if t.startswith("k1s1 %d->%d @%dx%d " % (Fr * E, E, S, S)))}
    k2_ms = sum(k2_parts.values())
    k2_survey_bytes = (2 * Fr + 2) * S * S * E * 2 * B        # what a gate pass + a fusion pass must move
    if "fbanet_faf_fuse_sm100" in fam and fam["fbanet_faf_fuse_sm100"][0] > 0:
        # one pass (ops.faf_fuse): features read once, fused map written once -- (F+1) H W E s; the survey's figure (which counts the
        # gated tensor written and re-read) is reported beside it
        t_ms, n, by = fam["fbanet_faf_fuse_sm100"]
        k2_min = (Fr + 1) * S * S * E * 2 * B
        hbm_kernels.append({"kernel": "faf_k2 = fbanet_faf_fuse_sm100 (scores + gates + 1x1 K=896 fusion + PReLU in one pass)", "bound": "hbm", "launches": n,
                            "ms": t_ms, "algorithmic_bytes": k2_min, "achieved": k2_min / t_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                            "frac": k2_min / t_ms / 1e6 / hbm_peak, "survey_8d_bytes": k2_survey_bytes,
                            "survey_8d_frac": k2_survey_bytes / t_ms / 1e6 / hbm_peak})
    elif k2_ms > 0:
        hbm_kernels.append({"kernel": "faf_k2 (score conv + gate apply + 1x1 K=896 fusion: three launches)", "bound": "hbm", "launches": 3, "ms": k2_ms,
                            "parts_ms": k2_parts, "algorithmic_bytes": k2_survey_bytes, "achieved": k2_survey_bytes / k2_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                            "frac": k2_survey_bytes / k2_ms / 1e6 / hbm_peak})
    # kernels bound by neither HBM nor the tensor pipe: what bounds them (ncu: profiles/r1_ncu_attn_dh16.txt, r2_ncu_leff_mlp_dec1.txt) and
    # their algorithmic bytes / time for reference
    cuda_core_kernels = []
    if "fbanet_window_attention_sm100" in fam:
        t_ms, n, by = fam["fbanet_window_attention_sm100"]
        exps = 1.4848e8 * B                      # softmax exponentials of one step (heads x 100 x 100 per window, all 20 layers)
        xu_peak = 16.0 * 148 * (pk.get("sm_max_mhz", 1965.0) * 1e6)
        cuda_core_kernels.append({"kernel": "fbanet_window_attention_sm100", "bound": "xu (MUFU.EX2, 16 / clk / SM)", "launches": n, "ms": t_ms,
                                  "achieved": exps / (t_ms * 1e-3) / 1e9, "peak": xu_peak / 1e9, "unit": "Gexp/s", "frac": exps / (t_ms * 1e-3) / xu_peak,
                                  "hbm_gbs": by / t_ms / 1e6, "hbm_frac": by / t_ms / 1e6 / hbm_peak})
    for name in ("fbanet_leff_mlp_sm100", "fbanet_leff_fc2_sm100"):
        if name in fam and fam[name][0] > 0:
            t_ms, n, by = fam[name]
            cuda_core_kernels.append({"kernel": name, "bound": "CUDA-core pipes (GELU: MUFU.TANH + packed FMA; depthwise 3x3: FFMA2)", "launches": n, "ms": t_ms,
                                      "hbm_gbs": by / t_ms / 1e6, "hbm_frac": by / t_ms / 1e6 / hbm_peak})
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, ts = cpu_forward_rate(threads, reps=5)
        cpu = {"value": v, "unit": "bursts/s", "cores": threads, "kind": "port",
               "sample": f"5 single-burst forwards (1/{B} of one step) of the CPU oracle, torch fp32, {threads} threads, median"}
    line = {
        "metric": "bursts_per_sec", "value": value, "unit": "bursts/s", "output_mp_per_s": value * MP_PER_BURST,
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype if args.dtype != "fp32" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"burst-sharded x{world}, no collective",
                   "l2": f"inputs {host_in.numel() * 4 / 1e6:.0f} MB/step + GB-scale activations > 126 MB L2", "cuda_graph": used_graph,
                   "weights": "random init (reference distributions), seed 0"},
        "e2e": {"value": e2e, "unit": "bursts/s", "h2d_bytes_per_step": host_in.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
                "ms_per_step": ms_e2e / args.steps, "api": "BaseModel.infer_host_stream (depth 2: one full-batch forward per step, the copies of "
                "neighbouring steps overlap it; every step's upload and download lie inside the timed region)"},
        "e2e_blocking": {"value": total_bursts / (ms_e2e_blocking / 1e3), "unit": "bursts/s", "ms_per_step": ms_e2e_blocking / args.steps,
                         "host_chunk": model.host_chunk, "api": "one blocking BaseModel.infer_host call per step (chunked 3-stream pipeline inside the call)"},
        "e2e_narrow_io": {"value": total_bursts / (ms_e2e_u8 / 1e3), "unit": "bursts/s", "in_dtype": "u8", "out_dtype": "u8",
                          "h2d_bytes_per_step": host_in.numel(), "d2h_bytes_per_step": host_out.numel(), "ms_per_step": ms_e2e_u8 / args.steps,
                          "note": "the reference's own 8-bit data path (uint8 frames / 255 in, clamp * 255 truncated out), conversions on the device; not the headline"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof, "roofline_classes": roofline_classes, "gemm1x1_shapes": gemm1x1_shapes, "conv_shapes": conv_shapes,
        "hbm_kernels": hbm_kernels, "cuda_core_kernels": cuda_core_kernels, "cpu_baseline": cpu, "clocks": clocks, "other_configs": others,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
