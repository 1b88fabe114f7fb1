"""World-size-2 ``gloo`` checks of the multi-rank plumbing (runs on CPU): rendezvous on 127.0.0.1, disjoint
burst shards, max-over-ranks timing, output gather.  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fbanet_b200.dist import gather_rows, init_from_env, reduce_max, shard_range
    r, _, w = init_from_env("gloo")
    n = 7
    b, e = shard_range(n, r, w)
    mine = torch.arange(n, dtype=torch.float32)[b:e].view(-1, 1) * 10 + r  # rank-tagged rows
    counts = [shard_range(n, i, w)[1] - shard_range(n, i, w)[0] for i in range(w)]
    full = gather_rows(mine, counts)
    slow = reduce_max(100.0 + 50.0 * r)  # rank 1 is "slower": the reported time is the max
    dist.barrier()
    q.put((r, (b, e), full.flatten().tolist(), slow))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, full0, t0), (r1, s1, full1, t1) = res
    assert s0 == (0, 4) and s1 == (4, 7)
    assert full0 == full1 == [0.0, 10.0, 20.0, 30.0, 41.0, 51.0, 61.0]
    assert t0 == t1 == 150.0


def _band_worker(rank, world, port, q):
    """cfg4 row bands on two ranks: each rank owns one band of the burst and a contiguous tile shard.  The halo rows it needs
    from the other band are fetched (here with gloo send/recv standing in for the NVLink peer loads of the CUDA kernel) exactly
    as `halo_sources` says, and the tiles it builds from (own band + halo) must equal the oracle's tiles of the whole image."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fbanet_b200.dist import band_rows, halo_sources, init_from_env, shard_range
    from oracle.fbanet_oracle import tensor_divide_burst
    r, _, w = init_from_env("gloo")
    T, C, H, W, ps, ov = 2, 3, 50, 30, 20, 10
    full = torch.rand(1, T, C, H, W, generator=torch.Generator().manual_seed(0))   # same on both ranks (seeded); only the band is "owned"
    row0 = band_rows(H, w)
    band = full[0, :, :, row0[r]:row0[r + 1]].clone()
    nh, nw = -(-H // ps), -(-W // ps)
    t0, t1 = shard_range(nh * nw, r, w)
    need = halo_sources(H, ps, ov, (t0 // nw, (t1 - 1) // nw + 1), row0)
    # exchange: every rank sends its whole band to whoever needs rows of it (a superset of the halo; sizes are asserted below)
    other = 1 - r
    theirs = torch.empty(T, C, row0[other + 1] - row0[other], W)
    reqs = [dist.isend(band, other), dist.irecv(theirs, other)]
    for x in reqs:
        x.wait()
    canvas = torch.full((T, C, H, W), float("nan"))
    canvas[:, :, row0[r]:row0[r + 1]] = band
    if other in need:
        canvas[:, :, row0[other]:row0[other + 1]] = theirs
    ref = tensor_divide_burst(full, ps, ov)[t0:t1]
    got = tensor_divide_burst(canvas[None], ps, ov)[t0:t1]      # NaN anywhere = a row `halo_sources` did not announce
    ok = bool(torch.equal(got, ref))
    dist.barrier()
    q.put((r, (t0, t1), dict(need), ok))
    dist.destroy_process_group()


def test_two_rank_row_bands_and_halo_plan():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, need0, ok0), (_, s1, need1, ok1) = res
    assert s0 == (0, 3) and s1 == (3, 6)                      # 3 x 2 tiles of 20 px over a 50 x 30 image
    assert ok0 and ok1
    assert need0 == {0: 25, 1: 25} and need1 == {0: 15, 1: 25}   # rank 0's tile rows 0-1 reach rows 0..49, rank 1's rows 10..49


def _flat_worker(rank, world, port, q):
    """Config 5's only collective: every rank's flat gradient buffer is summed in ONE all-reduce; the 1/world scale is handed to the
    optimizer step.  Parameters keep living inside the flat buffer (views), so state_dict / checkpoints are unaffected."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fbanet_b200.dist import init_from_env
    from fbanet_b200.train import FlatParams
    r, _, w = init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.PReLU(), torch.nn.Linear(7, 3))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(net.parameters())
    same = all(torch.equal(before[k], v) for k, v in net.state_dict().items())            # values survive the re-pointing
    views = all(p.data.data_ptr() >= flat.data.data_ptr() and p.data.data_ptr() < flat.data.data_ptr() + flat.numel * 4 for p in net.parameters())
    x = torch.full((4, 5), float(r + 1))
    net(x).sum().backward()                                                               # autograd accumulates INTO the flat gradient buffer
    local = flat.grad.clone()
    scale = flat.all_reduce()
    gathered = [torch.empty_like(local) for _ in range(w)]
    dist.all_gather(gathered, local)
    ok = bool(torch.allclose(flat.grad, sum(gathered))) and flat.grad.abs().sum().item() > 0
    dist.barrier()
    q.put((r, same, views, ok, scale, flat.numel))
    dist.destroy_process_group()


def test_two_rank_flat_gradient_all_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, same, views, ok, scale, n in res:
        assert same and views and ok and scale == 0.5 and n == 36 + 8 + 4 + 24 + 4          # each parameter padded to 16 bytes


def _bucket_worker(rank, world, port, q):
    """The bucketed form of config 5's collective: parameters marked ready in backward order (and one out of order), buckets launched
    in bucket order on every rank, result identical to the one-shot all-reduce."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from fbanet_b200.dist import init_from_env
    from fbanet_b200.train import FlatParams
    r, _, w = init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(16, 16) for _ in range(6)])
    flat = FlatParams(net.parameters())
    g = torch.Generator().manual_seed(100 + r)
    flat.grad.copy_(torch.randn(flat.numel, generator=g))
    local = flat.grad.clone()
    flat.begin_reduce(bucket_bytes=2 * (16 * 16 + 16) * 4)                                 # two layers per bucket -> 3 buckets
    nb, launched = len(flat._buckets), []
    order = list(reversed(flat.params))
    order[0], order[5] = order[5], order[0]                                                # one parameter becomes ready "late"
    for p in order:
        flat.mark_ready(p)
        launched.append(sum(flat._launched))
    scale = flat.finish_reduce()
    bucketed = flat.grad.clone()
    flat.grad.copy_(local)
    flat.all_reduce()
    covered = sorted((b, e) for b, e, _ in flat._buckets)
    tiles = covered[0][0] == 0 and covered[-1][1] == flat.numel and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    flat.begin_reduce(bucket_bytes=2 * (16 * 16 + 16) * 4)                                 # second step: nothing marked, finish launches all
    flat.grad.copy_(local)
    flat.finish_reduce()
    again = torch.equal(flat.grad, bucketed)
    dist.barrier()
    q.put((r, nb, launched, scale, torch.equal(bucketed, flat.grad) and again, tiles))
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_reduce_equals_one_shot():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, nb, launched, scale, same, tiles in res:
        assert nb == 3 and scale == 0.5 and same and tiles
        # the tail bucket (layers 5, 4) needs the late parameter (layer 5's bias, marked 6th): nothing launches before it; the middle
        # bucket (layers 3, 2) completes with layer 2's weight (8th), the head bucket with the last parameter
        assert launched[:5] == [0] * 5 and launched[5:7] == [1, 1] and launched[7] == 2 and launched[-2:] == [2, 3], launched


def _train_step_worker(rank, world, port, q):
    """Config 5 data-parallel: each rank runs train.train_step on ITS burst (op stand-ins from tests/test_host_logic.py, the only
    collective is the bucketed sum of the flat gradient buffer); the summed gradient times 1 / world equals the gradient of one
    process over the two-burst batch, and both ranks end the step with identical parameters."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_host_logic as H
    from fbanet_b200 import ops, train
    from fbanet_b200.dist import init_from_env
    from fbanet_b200.model import BaseModel

    class MP:
        def setattr(self, o, n, v):
            setattr(o, n, v)
    H._install_op_standins(MP())
    H._install_conv_standins(MP())

    def adam_step(param, grad, m, v, step, lr, betas, eps, wd, decoupled, grad_scale):
        g = grad * grad_scale
        param.mul_(1 - lr * wd)
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        param.addcdiv_(m / (1 - betas[0] ** step), (v / (1 - betas[1] ** step)).sqrt() + eps, value=-lr)
    ops.adam_step = adam_step
    r, _, w = init_from_env("gloo")
    torch.set_num_threads(2)

    def make():
        m = BaseModel(token_mlp="leff", dtype="fp32", seed=3, num_frames=2, img_size=16, embed_dim=16, window_length=4)
        m.drop_path_rate = 0.0
        for p in m.parameters():
            p.requires_grad_(True)
        return m, train.FlatParams(m.parameters())
    g = torch.Generator().manual_seed(9)
    bursts, targets = torch.rand(w, 2, 3, 16, 16, generator=g), torch.rand(w, 3, 64, 64, generator=g)
    m, flat = make()
    train.train_step(m, flat, bursts[r:r + 1], targets[r:r + 1], lr=1e-3, bucket_bytes=64 << 10)
    dp_grad = flat.grad * (1.0 / w)
    both = [torch.empty_like(flat.data) for _ in range(w)]
    dist.all_gather(both, flat.data)
    same_params = all(torch.equal(both[0], b) for b in both)
    # one process, the whole batch (no process group involvement: world-1 semantics through a fresh FlatParams and all_reduce skipped)
    m1, flat1 = make()
    restored, tape = train.model_forward_train(m1, bursts, training=False)
    loss, d = ops.training_loss(restored, targets, clamp_restored=True)      # what train_step computes (train.py.bak:167)
    tape.backward(restored, d)
    err = ((dp_grad - flat1.grad).abs().max() / flat1.grad.abs().max()).item()
    dist.barrier()
    q.put((r, same_params, err, len(flat._buckets)))
    dist.destroy_process_group()


def test_two_rank_data_parallel_train_step():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_step_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r, same_params, err, nb in res:
        assert same_params and err < 1e-4 and nb > 10, (r, same_params, err, nb)
