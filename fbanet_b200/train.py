"""Training-step plumbing around the backward bricks -- SURVEY 8f-3, BASELINE config 5:
flat parameter / gradient / moment buffers, the data-parallel gradient all-reduce (the ONE collective of the whole system:
``ncclAllReduce(sum)`` over the 19.2 M gradients, reference ``train.py.bak:83-84`` ``DataParallel`` / north-star "NCCL allreduce appears only
in the training-step config"), the fused Adam / AdamW step (``train.py.bak:72-78``) and the loss (``:118-119,168``).

What a training step will be once the backward exists::

    loss, d_restored = ops.training_loss(model(burst), target)      # built (fbanet_train_loss_sm100)
    backward(model, d_restored) -> flat.grad                        # NOT built: DESIGN.md 8c
    flat.all_reduce()                                               # built: one NCCL sum over the flat gradient buffer (or begin_reduce /
                                                                    # mark_ready / finish_reduce: the same sum in buckets, overlapped)
    flat.adam_step(lr)                                              # built (fbanet_adam_step_sm100), grad_scale = 1 / world
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatParams:
    """All parameters of a module re-pointed into ONE contiguous fp32 buffer (plus flat gradient and Adam moment buffers of the same
    layout), so the all-reduce is one collective and the optimizer one bandwidth pass.  ``module.state_dict()`` / checkpoints are
    unaffected: the parameters keep their names and shapes, only their storage moves."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params), "one device, fp32 master parameters"
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.data = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = 0
        for p, o in zip(self.params, self.offsets):
            self.data[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + p.numel()].view_as(p)                # the parameter now lives inside the flat buffer
            p.grad = self.grad[o:o + p.numel()].view_as(p)                # and its gradient inside the flat gradient buffer

    def zero_grad(self) -> None:
        self.grad.zero_()

    def all_reduce(self, group=None) -> float:
        """Sum the flat gradient over the data-parallel ranks (one collective); returns the scale (1 / world) the optimizer step
        must apply -- the division is fused into ``adam_step`` instead of being a pass of its own."""
        if not (dist.is_available() and dist.is_initialized()):
            return 1.0
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        return 1.0 / world

    # ---- the same sum as bucketed collectives issued WHILE the backward still runs -------------------------
    def begin_reduce(self, group=None, bucket_bytes: int = 16 << 20) -> None:
        """Start a bucketed reduction of the flat gradient: the parameters are grouped, from the TAIL of the buffer (the backward
        reaches the last layers first), into runs of at least ``bucket_bytes``; :meth:`mark_ready` launches a run's
        ``all_reduce(sum, async_op=True)`` as soon as the gradient of every parameter in it is final, so the transfers overlap the
        rest of the backward; :meth:`finish_reduce` launches what is left and waits.  On NVSwitch the cost of a collective does not
        depend on link count, so the bucket size only trades launch latency (few, large) against overlap (many, small): 16 MiB
        gives 5 buckets for the 77 MB of fp32 gradients.  Every element takes part in exactly one collective, so the result equals
        :meth:`all_reduce`'s."""
        self._group = group
        self._world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if getattr(self, "_bucket_bytes", None) != bucket_bytes:
            self._bucket_bytes, self._buckets, self._bucket_of = bucket_bytes, [], {}
            end, size = self.numel, 0
            for k in range(len(self.params) - 1, -1, -1):
                size += self.params[k].numel() * 4
                if size >= bucket_bytes or k == 0:
                    self._buckets.append((self.offsets[k], end, k))                      # [begin, end) in elements, first param index
                    end, size = self.offsets[k], 0
            first = [b[2] for b in self._buckets]
            for bi, f in enumerate(first):
                last = first[bi - 1] if bi > 0 else len(self.params)
                for k in range(f, last):
                    self._bucket_of[id(self.params[k])] = bi
        self._pending = [0] * len(self._buckets)
        for p in self.params:
            self._pending[self._bucket_of[id(p)]] += 1
        self._ready, self._launched, self._handles = set(), [False] * len(self._buckets), []

    def _launch(self, bi: int) -> None:
        if self._launched[bi]:
            return
        self._launched[bi] = True
        if self._world > 1:
            b, e, _ = self._buckets[bi]
            self._handles.append(dist.all_reduce(self.grad[b:e], op=dist.ReduceOp.SUM, group=self._group, async_op=True))

    def mark_ready(self, param: torch.nn.Parameter) -> None:
        """The gradient of ``param`` is final (the backward will not touch it again in this step).  Buckets are launched in bucket
        order only -- every rank must issue the same sequence of collectives -- so a bucket that completes early waits for the ones
        before it."""
        if id(param) in self._ready:
            return
        self._ready.add(id(param))
        self._pending[self._bucket_of[id(param)]] -= 1
        for bi in range(len(self._buckets)):
            if self._launched[bi]:
                continue
            if self._pending[bi] > 0:
                break
            self._launch(bi)

    def finish_reduce(self) -> float:
        """Launch the buckets not yet launched, wait for all of them; returns 1 / world like :meth:`all_reduce`."""
        for bi in range(len(self._buckets)):
            self._launch(bi)
        for h in self._handles:
            h.wait()
        self._handles = []
        return 1.0 / self._world

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, decoupled: bool = True,
                  grad_scale: float = 1.0) -> None:
        """One fused Adam / AdamW update of every parameter (``fbanet_adam_step_sm100``); CUDA only, no fallback."""
        from . import ops
        self.step += 1
        ops.adam_step(self.data, self.grad, self.exp_avg, self.exp_avg_sq, self.step, lr, betas, eps, weight_decay, decoupled, grad_scale)

    def optimizer_state(self) -> dict:
        return {"step": self.step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}


def dgrad_weight(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Packed weight that turns the FORWARD implicit GEMM into the data gradient of a stride-1 convolution / linear layer:
    for ``y = conv(x, w)`` with ``w [Co,Ci,kh,kw]``, stride 1 and padding ``pad``, ``dx = conv(dy, w')`` with ``w'[ci][ky][kx][co] =
    w[co][ci][kh-1-ky][kw-1-kx]`` and padding ``k - 1 - pad`` -- i.e. ``ops.conv_gemm([dy], dgrad_weight(w, dt), dx, kh=kh, kw=kw,
    pad=kh-1-pad)``; a linear layer ``[out, in]`` is the ``kh = kw = 1`` case (``dx = dy @ w``).  Returned in ``conv_gemm``'s packed
    layout ``[Ci, kh*kw*Co]`` (K index ``(ky*kw + kx)*Co + co``), so in bf16 the data gradients run on the tcgen05 kernel."""
    w = weight.detach()
    if w.dim() == 2:
        w = w[:, :, None, None]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------------------------------------
# learning-rate schedules and stochastic-depth rates of the training configuration (host arithmetic, no tensors)
# ------------------------------------------------------------------------------------------------------------------------------
def warmup_cosine_lr(epoch: int, lr_initial: float = 1e-4, nepoch: int = 250, warmup_epochs: int = 3, eta_min: float = 1e-6) -> float:
    """Learning rate DURING training epoch ``epoch`` (1-based, as the loop of ``train.py.bak:150`` counts) under ``--warmup``:
    ``GradualWarmupScheduler(multiplier=1, total_epoch=warmup_epochs, after_scheduler=CosineAnnealingLR(nepoch - warmup_epochs,
    eta_min=1e-6))`` stepped once before the first epoch and once after every epoch (``train.py.bak:104-110,220``,
    ``warmup_scheduler/scheduler.py:24-39,57-67``).  Closed form of what that pair of objects actually does (checked against the
    reference's scheduler executed over whole runs, ``tests/golden/lr_schedule_reference.npz``): a linear ramp ``lr * e / warmup`` for
    ``e <= warmup`` (epoch ``warmup`` already runs at the full rate); at the hand-over the wrapper returns the cosine scheduler's
    RECURSIVE ``get_lr()`` while that scheduler still sits at its epoch 0, so with ``c = e - warmup - 1`` and ``T = nepoch - warmup``
    the rate is ``eta + (lr - eta) * (1 + cos(pi c / T)) / (1 + cos(pi / T))``: a slight overshoot above ``lr`` at ``c = 0``,
    exactly ``lr`` at ``c = 1``, and a cosine that ends at ``c = T - 1`` (one epoch short of ``eta``)."""
    import math
    if epoch <= warmup_epochs:
        return lr_initial * epoch / warmup_epochs
    c, t_max = epoch - warmup_epochs - 1, nepoch - warmup_epochs
    return eta_min + (lr_initial - eta_min) * (1.0 + math.cos(math.pi * c / t_max)) / (1.0 + math.cos(math.pi / t_max))


def step_lr(epoch: int, lr_initial: float = 1e-4, step: int = 50, gamma: float = 0.5) -> float:
    """Learning rate during epoch ``epoch`` (1-based) without ``--warmup``: ``StepLR(step_size=50, gamma=0.5)`` stepped once before
    the first epoch (``train.py.bak:111-115``)."""
    return lr_initial * gamma ** (epoch // step)


def drop_path_rates(depths=(2, 2, 2, 2, 2, 2, 2, 2, 2), drop_path_rate: float = 0.1) -> dict:
    """Per-layer stochastic-depth rates as ``models/fba_net.py:96-100`` hands them to the blocks of BOTH hourglasses:
    ``enc = linspace(0, rate, sum(depths[:4]))``, bottleneck ``[rate] * depths[4]``, ``dec = enc[::-1]``; encoder 0 takes
    ``enc[:d0]``, encoder 1 ``enc[d0:d0+d1]``, decoder 0 ``dec[:d5]``, decoder 1 ``dec[d5:d5+d6]`` (``:130-229``)."""
    n = sum(depths[: len(depths) // 2])
    enc = [drop_path_rate * i / (n - 1) for i in range(n)] if n > 1 else [0.0]
    dec = enc[::-1]
    d0, d1, d4, d5, d6 = depths[0], depths[1], depths[4], depths[5], depths[6]
    return {"encoderlayer_0": enc[:d0], "encoderlayer_1": enc[d0:d0 + d1], "conv": [drop_path_rate] * d4,
            "decoderlayer_0": dec[:d5], "decoderlayer_1": dec[d5:d5 + d6]}


def drop_path_scales(batch: int, rate: float, generator: Optional[torch.Generator] = None, device=None) -> torch.Tensor:
    """Per-burst multipliers of one ``DropPath`` call (``layers/drop_path.py:52-63``, "global" mode mapped over the batch by
    ``jax.vmap``): ``bernoulli(keep) / keep`` with ``keep = 1 - rate``; all ones for ``rate == 0`` (the layer returns ``x`` itself),
    all zeros for ``rate == 1`` (the reference skips the division when ``keep == 0``).  fp32 ``[batch]`` for
    ``ops.drop_path_add``; drawn on the host generator so that every data-parallel rank can reproduce its own stream."""
    assert 0.0 <= rate <= 1.0
    keep = 1.0 - rate
    if rate == 0.0:
        s = torch.ones(batch, dtype=torch.float32)
    else:
        s = torch.bernoulli(torch.full((batch,), keep, dtype=torch.float32), generator=generator)
        if keep > 0.0:
            s = s / keep
    return s.to(device) if device is not None else s
