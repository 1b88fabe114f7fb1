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
            n += (p.numel() + 3) // 4 * 4          # every parameter starts on a 16-byte boundary (vector loads of biases / tables)
        self.numel = n                              # padding elements stay zero: zero gradient, zero update
        self.data = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = 0
        self.generation = 0                         # bumped by every raw-pointer update of `data` (see adam_step)
        for p, o in zip(self.params, self.offsets):
            self.data[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + p.numel()].view_as(p)                # the parameter now lives inside the flat buffer
            p.grad = self.grad[o:o + p.numel()].view_as(p)                # and its gradient inside the flat gradient buffer
            p._fbanet_flat = self                                         # BaseModel._signature reads `generation` through this

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
        if id(param) in self._ready or id(param) not in self._bucket_of:      # (a frozen parameter is not part of the flat buffers)
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
        ev = getattr(self, "reduce_events", None)          # a list: record (before, after) CUDA events = the EXPOSED reduction time
        if ev is not None and self.grad.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for bi in range(len(self._buckets)):
            self._launch(bi)
        for h in self._handles:
            h.wait()
        self._handles = []
        if ev is not None and self.grad.is_cuda:
            e1.record()
            ev.append((e0, e1))
        return 1.0 / self._world

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, decoupled: bool = True,
                  grad_scale: float = 1.0) -> None:
        """One fused Adam / AdamW update of every parameter (``fbanet_adam_step_sm100``); CUDA only, no fallback."""
        from . import ops
        self.step += 1
        self.param_groups[0].update(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        ops.adam_step(self.data, self.grad, self.exp_avg, self.exp_avg_sq, self.step, lr, betas, eps, weight_decay, decoupled, grad_scale)
        # the kernel writes through a raw pointer: neither a parameter's version counter nor its address changes, so the model's
        # packed-weight / CUDA-graph caches (BaseModel.packed, _host_graph) are told explicitly that the weights moved
        self.generation += 1

    def optimizer_state(self) -> dict:
        return {"step": self.step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    # ---- checkpoint / resume in the reference's format: ``{"optimizer": optimizer.state_dict()}`` of ``optim.Adam / AdamW``
    # (``train.py.bak:72-78,199-246``; read back by ``utils/model_utils.py:51-62`` ``load_optim(optimizer, weights)``, which only needs
    # ``load_state_dict`` and ``param_groups``) -- so a FlatParams can be handed to the reference's helpers in the optimizer's place,
    # and checkpoints written by either side resume on the other.
    @property
    def param_groups(self):
        if not hasattr(self, "hyper"):
            self.hyper = {"lr": 1e-4, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0.0, "amsgrad": False}
        return [self.hyper]

    def state_dict(self) -> dict:
        state = {}
        if self.step > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step)), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0], params=list(range(len(self.params))))]}

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        index = [i for g in groups for i in g["params"]]
        assert len(index) == len(self.params), f"optimizer state holds {len(index)} parameters, the model has {len(self.params)}"
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for k, (p, o) in zip(index, zip(self.params, self.offsets)):
            st = sd["state"].get(k)
            if st is None:
                continue
            n = p.numel()
            assert st["exp_avg"].numel() == n, f"parameter {k}: state of {st['exp_avg'].numel()} elements for {n}"
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        assert len(steps) <= 1, f"parameters at different step counts {sorted(steps)}: not a state this fused step can continue"
        self.step = steps.pop() if steps else 0
        self.hyper = {k: v for k, v in groups[0].items() if k != "params"}


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


def dgrad_weight_convT2(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Data gradient of ``ConvTranspose2d(Ci, Co, 2, 2)`` (``layers/upsample.py:19-30``, ``upsample_flatten.py:6-12``; weight
    ``[Ci,Co,2,2]``) as a 1x1 GEMM over the space-to-depth view of ``dy``: ``dx(y,x,ci) = sum_{a,b,co} dy(2y+a, 2x+b, co) w[ci,co,a,b]``
    = ``ops.conv_gemm([ops.space_to_depth(dy)], dgrad_weight_convT2(w, dt), dx)``.  Packed ``[Ci, 4*Co]``, K index ``(a*2+b)*Co + co``
    (the channel order ``fbanet_space_to_depth_sm100`` writes)."""
    w = weight.detach()
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(dtype).contiguous()


def dgrad_weight_down4(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Data gradient of ``Conv2d(Ci, Co, 4, stride 2, padding 1)`` (``layers/downsample.py:19-30``, ``downsample_flatten.py:6-13``;
    weight ``[Co,Ci,4,4]``) as ONE 3x3 implicit GEMM over ``dy`` whose 4*Ci output rows are the four sub-pixel phases of ``dx``,
    scattered by the forward's ConvTranspose-style store: input pixel ``(2i+a, 2j+b)`` is read by output pixel ``i + d`` through
    tap ``ky = a + 1 - 2d`` (``d`` in {-1,0,1}, two of the three valid per phase), so
    ``W'[(a*2+b)*Ci + ci][((d+1)*3 + (e+1))*Co + co] = w[co,ci,a+1-2d,b+1-2e]`` (0 where the tap falls outside 0..3) and
    ``dx = ops.conv_gemm([dy], W', dx[N,2Ho,2Wo,Ci], kh=3, kw=3, pad=1, store_mode=STORE_CONVT2)``.  (5/9 of the MACs multiply
    zeros; the kernel is the verified forward one.)"""
    w = weight.detach()
    Co, Ci = w.shape[:2]
    out = torch.zeros(2, 2, Ci, 3, 3, Co, dtype=w.dtype, device=w.device)
    for a in range(2):
        for ty in range(3):
            ky = a + 3 - 2 * ty
            if not 0 <= ky < 4:
                continue
            for b in range(2):
                for tx in range(3):
                    kx = b + 3 - 2 * tx
                    if 0 <= kx < 4:
                        out[a, b, :, ty, tx, :] = w[:, :, ky, kx].t()
    return out.reshape(4 * Ci, 9 * Co).to(dtype).contiguous()


def dgrad_weight_pixel_shuffle(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Data gradient of ``Conv2d(Ci, 4C, 3, padding 1)`` followed by ``PixelShuffle(2)`` (``blocks/upsampler.py:22-32``,
    ``layers/pixel_shuffle.py:9-10``; weight ``[4C,Ci,3,3]``): the un-shuffled gradient ``dz(y,x,4c+2i+j) = dy(2y+i,2x+j,c)`` is the
    space-to-depth view of ``dy`` with its channels in ``(2i+j)*C + c`` order, so the permutation goes into the K axis of the
    ordinary flipped-weight packing: ``dx = ops.conv_gemm([ops.space_to_depth(dy)], W', dx, kh=3, kw=3, pad=1)``."""
    w = weight.detach()
    C4, Ci = w.shape[:2]
    wf = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, 3, 3, C4 // 4, 4).permute(0, 1, 2, 4, 3)     # [ci, ky, kx, (i,j), c]
    return wf.reshape(Ci, -1).to(dtype).contiguous()


def convT2_dgrad(dy: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``dy [N,2H,2W,Co]`` channels-last -> ``dx [N,H,W,Ci]`` for ``ConvTranspose2d(Ci,Co,2,2)`` (CUDA only)."""
    from . import ops
    N, H2, W2, _ = dy.shape
    dx = torch.empty((N, H2 // 2, W2 // 2, weight.shape[0]), device=dy.device, dtype=dy.dtype)
    return ops.conv_gemm([ops.space_to_depth(dy)], dgrad_weight_convT2(weight, dy.dtype), dx)


def down4_dgrad(dy: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``dy [N,Ho,Wo,Co]`` channels-last -> ``dx [N,2Ho,2Wo,Ci]`` for ``Conv2d(Ci,Co,4,2,1)`` (CUDA only)."""
    from . import ops, _lib as L
    N, Ho, Wo, _ = dy.shape
    dx = torch.empty((N, 2 * Ho, 2 * Wo, weight.shape[1]), device=dy.device, dtype=dy.dtype)
    return ops.conv_gemm([dy], dgrad_weight_down4(weight, dy.dtype), dx, kh=3, kw=3, pad=1, store_mode=L.STORE_CONVT2)


def pixel_shuffle_conv_dgrad(dy: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``dy [N,2H,2W,C]`` channels-last (gradient of the SHUFFLED output) -> ``dx [N,H,W,Ci]`` for conv3x3 ``[4C,Ci,3,3]`` +
    ``PixelShuffle(2)`` (CUDA only)."""
    from . import ops
    N, H2, W2, _ = dy.shape
    dx = torch.empty((N, H2 // 2, W2 // 2, weight.shape[1]), device=dy.device, dtype=dy.dtype)
    return ops.conv_gemm([ops.space_to_depth(dy)], dgrad_weight_pixel_shuffle(weight, dy.dtype), dx, kh=3, kw=3, pad=1)


def convT2_wgrad(x: torch.Tensor, dy: torch.Tensor):
    """Weight / bias gradient of ``ConvTranspose2d(Ci,Co,2,2)`` in the torch layouts (``[Ci,Co,2,2]``, ``[Co]``): the 1x1 weight
    gradient between ``x [N,H,W,Ci]`` and the space-to-depth view of ``dy [N,2H,2W,Co]`` gives rows ``(a,b,co)``; the bias sums
    the four phases (CUDA only)."""
    from . import ops
    Co = dy.shape[-1]
    dw4, db4 = ops.conv_wgrad(x, ops.space_to_depth(dy))                                          # [4Co, Ci, 1, 1], [4Co]
    return convT2_wgrad_layout(dw4, db4, Co)


def convT2_wgrad_layout(dw4: torch.Tensor, db4: torch.Tensor, Co: int):
    Ci = dw4.shape[1]
    return dw4.reshape(2, 2, Co, Ci).permute(3, 2, 0, 1).contiguous(), db4.reshape(4, Co).sum(0)


def pixel_shuffle_conv_wgrad(x: torch.Tensor, dy: torch.Tensor):
    """Weight / bias gradient of conv3x3 ``[4C,Ci,3,3]`` + ``PixelShuffle(2)``: the 3x3 weight gradient between ``x`` and the
    space-to-depth view of ``dy`` has its output channels in ``(2i+j)*C + c`` order; re-ordered to torch's ``4c+2i+j`` (CUDA only)."""
    from . import ops
    dw, db = ops.conv_wgrad(x, ops.space_to_depth(dy), 3, 3, 1, 1)
    return pixel_shuffle_wgrad_layout(dw, db)


def pixel_shuffle_wgrad_layout(dw: torch.Tensor, db: torch.Tensor):
    C4 = dw.shape[0]
    return (dw.reshape(4, C4 // 4, *dw.shape[1:]).transpose(0, 1).reshape(dw.shape).contiguous(),
            db.reshape(4, C4 // 4).t().reshape(-1).contiguous())


# ------------------------------------------------------------------------------------------------------------------------------
# one LeWin block in training mode: forward that keeps its activations, backward composed of the bricks
# ------------------------------------------------------------------------------------------------------------------------------
_REDUCER: Optional["FlatParams"] = None      # set by train_step while a bucketed reduction is open


def _accumulate(p: torch.nn.Parameter, g: torch.Tensor) -> None:
    """Add ``g`` to ``p.grad`` (a view into ``FlatParams.grad`` when the module was flattened; created on first use otherwise).
    Every parameter of the model is consumed by exactly one tape record, so after this call its gradient is final for the step:
    with a bucketed reduction open, the parameter is reported ready and complete buckets start their all-reduce while the rest of
    the backward is still running."""
    g = g.reshape(p.shape).to(p.dtype)
    if p.grad is None:
        p.grad = g.clone()
    else:
        p.grad.add_(g)
    if _REDUCER is not None:
        _REDUCER.mark_ready(p)


def _linear(x4: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    from . import ops
    out = torch.empty((*x4.shape[:3], weight.shape[0]), device=x4.device, dtype=x4.dtype)
    return ops.conv_gemm([x4], weight.detach().to(x4.dtype).contiguous(), out, bias=None if bias is None else bias.detach().contiguous())


def _linear_backward(lin_weight: torch.Tensor, x4: torch.Tensor, dy4: torch.Tensor):
    """``(dx, dW [out,in], db)`` of ``y = x W^T + b``: data gradient through the forward GEMM, weight gradient through the wgrad kernel."""
    from . import ops
    dx = torch.empty_like(x4)
    ops.conv_gemm([dy4], dgrad_weight(lin_weight, dy4.dtype), dx)
    dw, db = ops.conv_wgrad(x4, dy4)
    return dx, dw.reshape(dw.shape[0], dw.shape[1]), db


def lewin_forward_train(ly, x: torch.Tensor, s_attn: Optional[torch.Tensor] = None, s_mlp: Optional[torch.Tensor] = None,
                        gelu_act: Optional[int] = None, eps: float = 1e-5, qk_scale: Optional[float] = None):
    """(Parameters are used in their stored dtype: the fp32 masters the C-ABI ops require.)
    Training-mode forward of one LeWin block (``FBANetLayer.__call__``, ``layers/fba_net.py:139-250`` with the residuals of SURVEY
    Appendix A-4): ``x1 = x + s_attn * proj(attn(qkv(LN1 x)))``, ``y = x1 + s_mlp * fc2(gelu(dw(gelu(fc1(LN2 x1)))))``.  ``ly``: the
    model's layer module (``model._Layer``: norm1, attn.{qkv.to_q, qkv.to_kv, proj, relative_position_bias_table}, norm2,
    mlp.{linear1, dwconv, linear2}); ``x`` ``[B,H,W,C]`` channels-last on the GPU (fp32 or bf16); ``s_attn`` / ``s_mlp``: per-burst
    DropPath multipliers (:func:`drop_path_scales`; ``None`` = ones, the layer's rate is 0 or the model is in eval mode).
    Returns ``(y, saved)``; every step is one of the C-ABI ops, the activations ``saved`` keeps are what :func:`lewin_backward` reads."""
    from . import ops, _lib as L
    gelu_act = L.ACT_GELU_TANH if gelu_act is None else gelu_act
    B, H, W, C = x.shape
    T = B * H * W
    ones = torch.ones(B, device=x.device, dtype=torch.float32)
    s_attn = ones if s_attn is None else s_attn
    s_mlp = ones if s_mlp is None else s_mlp
    a = ly.attn
    scale = qk_scale or (C // ly.heads) ** -0.5          # `qk_scale or head_dim ** -0.5` (layers/window_attention.py:141-142)
    ln1 = ops.layernorm(x.view(T, C), ly.norm1.weight.detach(), ly.norm1.bias.detach(), eps).view(B, H, W, C)
    wqkv = torch.cat([a.qkv.to_q.weight.detach(), a.qkv.to_kv.weight.detach()], 0)
    bqkv = torch.cat([a.qkv.to_q.bias.detach(), a.qkv.to_kv.bias.detach()], 0)
    qkv = _linear(ln1, wqkv, bqkv)
    rpb = a.relative_position_bias_table.detach().contiguous()
    att = ops.window_attention(qkv.view(T, 3 * C), rpb, B, H, W, ly.heads, ly.win, ly.shift, scale).view(B, H, W, C)
    pr = _linear(att, a.proj.weight, a.proj.bias)
    x1 = ops.drop_path_add(pr, s_attn, skip=x)
    ln2 = ops.layernorm(x1.view(T, C), ly.norm2.weight.detach(), ly.norm2.bias.detach(), eps).view(B, H, W, C)
    fc1, dw, fc2 = ly.mlp.linear1[0], ly.mlp.dwconv[0], ly.mlp.linear2[0]
    h0 = _linear(ln2, fc1.weight, fc1.bias)                                     # pre-activations are kept for the backward
    h1 = ops.act_forward(h0, gelu_act)
    w9c = dw.weight.detach().reshape(dw.weight.shape[0], 9).t().contiguous()
    d0 = ops.dwconv3x3(h1, w9c, dw.bias.detach().contiguous(), L.ACT_NONE)
    d1 = ops.act_forward(d0, gelu_act)
    m = _linear(d1, fc2.weight, fc2.bias)
    y = ops.drop_path_add(m, s_mlp, skip=x1)
    saved = dict(x=x, ln1=ln1, qkv=qkv, att=att, x1=x1, ln2=ln2, h0=h0, h1=h1, d0=d0, d1=d1, s_attn=s_attn, s_mlp=s_mlp, wqkv=wqkv,
                 w9c=w9c, rpb=rpb, scale=scale, gelu_act=gelu_act, eps=eps, ones=ones)
    return y, saved


def lewin_backward(ly, saved: dict, dy: torch.Tensor) -> torch.Tensor:
    """Backward of :func:`lewin_forward_train`: returns ``dx`` and ACCUMULATES the gradients of the layer's 17 parameter tensors into
    their ``.grad`` (views of ``FlatParams.grad`` when flattened).  Composition of the bricks only: ``fbanet_drop_path_add`` (branch
    scaling and the residual sums), ``fbanet_conv_gemm`` with :func:`dgrad_weight` (data gradients of the four linear layers),
    ``fbanet_wgrad``, ``fbanet_act_bwd``, ``fbanet_dwconv3x3_bwd``, ``fbanet_layernorm_bwd``, ``fbanet_window_attention_bwd``."""
    from . import ops
    S = saved
    B, H, W, C = S["x"].shape
    T = B * H * W
    a, fc1, dw, fc2 = ly.attn, ly.mlp.linear1[0], ly.mlp.dwconv[0], ly.mlp.linear2[0]
    act = S["gelu_act"]
    # ---- LeFF branch: y = x1 + s_mlp * fc2(d1)
    dm = ops.drop_path_add(dy.contiguous(), S["s_mlp"])
    dd1, g_w, g_b = _linear_backward(fc2.weight, S["d1"], dm)
    _accumulate(fc2.weight, g_w), _accumulate(fc2.bias, g_b)
    dd0 = ops.act_backward(S["d0"], dd1, act)
    dh1, g_w, g_b = ops.dwconv3x3_backward(S["h1"], dd0, S["w9c"])
    _accumulate(dw.weight, g_w), _accumulate(dw.bias, g_b)
    dh0 = ops.act_backward(S["h0"], dh1, act)
    dln2, g_w, g_b = _linear_backward(fc1.weight, S["ln2"], dh0)
    _accumulate(fc1.weight, g_w), _accumulate(fc1.bias, g_b)
    dx1n, g_g, g_b = ops.layernorm_backward(S["x1"].view(T, C), dln2.view(T, C), ly.norm2.weight.detach().contiguous(), S["eps"])
    _accumulate(ly.norm2.weight, g_g), _accumulate(ly.norm2.bias, g_b)
    dx1 = ops.drop_path_add(dx1n.view(B, H, W, C), S["ones"], skip=dy.contiguous())             # skip path + LayerNorm path
    # ---- attention branch: x1 = x + s_attn * proj(att)
    dpr = ops.drop_path_add(dx1, S["s_attn"])
    datt, g_w, g_b = _linear_backward(a.proj.weight, S["att"], dpr)
    _accumulate(a.proj.weight, g_w), _accumulate(a.proj.bias, g_b)
    dqkv, g_t = ops.window_attention_backward(S["qkv"].view(T, 3 * C), datt.view(T, C), S["rpb"], B, H, W, ly.heads, ly.win, ly.shift, S["scale"])
    _accumulate(a.relative_position_bias_table, g_t)
    dln1, g_w, g_b = _linear_backward(S["wqkv"], S["ln1"], dqkv.view(B, H, W, 3 * C))
    _accumulate(a.qkv.to_q.weight, g_w[:C]), _accumulate(a.qkv.to_q.bias, g_b[:C])
    _accumulate(a.qkv.to_kv.weight, g_w[C:]), _accumulate(a.qkv.to_kv.bias, g_b[C:])
    dxn, g_g, g_b = ops.layernorm_backward(S["x"].view(T, C), dln1.view(T, C), ly.norm1.weight.detach().contiguous(), S["eps"])
    _accumulate(ly.norm1.weight, g_g), _accumulate(ly.norm1.bias, g_b)
    return ops.drop_path_add(dxn.view(B, H, W, C), S["ones"], skip=dx1)


# ------------------------------------------------------------------------------------------------------------------------------
# reverse-mode tape over the C-ABI ops, and the LeWin hourglass (models/fba_net.py:271-287) on it
# ------------------------------------------------------------------------------------------------------------------------------
class Tape:
    """Minimal reverse-mode tape: the training forward records, per composite op, its output tensor and a closure that maps the
    output's gradient to ``[(input tensor, gradient), ...]`` (parameter gradients are accumulated into ``.grad`` inside the closure).
    :meth:`backward` walks the records in reverse; a tensor consumed by several ops (skip connections) gets the SUM of its
    gradients, formed by ``fbanet_drop_path_add`` with unit scales.  No torch autograd, no torch arithmetic: torch only owns the
    buffers (``torch.cat`` / slicing of the concat inputs are copies)."""

    def __init__(self):
        self.records = []

    def record(self, out: torch.Tensor, backward_fn) -> torch.Tensor:
        self.records.append((out, backward_fn))
        return out

    def backward(self, out: torch.Tensor, dout: torch.Tensor) -> dict:
        from . import ops
        grads = {id(out): dout}
        for o, fn in reversed(self.records):
            g = grads.pop(id(o), None)
            if g is None:
                continue                                                   # this output did not reach the loss
            for t, gt in fn(g):
                gt = gt.contiguous()
                if id(t) in grads:
                    ones = torch.ones(gt.shape[0], device=gt.device, dtype=torch.float32)
                    grads[id(t)] = ops.drop_path_add(gt, ones, skip=grads[id(t)].contiguous())
                else:
                    grads[id(t)] = gt
        return grads                                                       # what is left: gradients of the tape's inputs, by id()


def _t_block(tape: Tape, block, x: torch.Tensor, rates, scales, training: bool, gelu_act: Optional[int] = None,
             qk_scale: Optional[float] = None) -> torch.Tensor:
    """A ``_Block`` (``blocks/fba_net.py:35-65``): its LeWin layers in sequence, each with its own stochastic-depth rate.
    ``scales``: the step's :class:`DropPathDraws` (per-burst multipliers of every DropPath call, drawn once per step)."""
    for ly, rate in zip(block.blocks, rates):
        B = x.shape[0]
        s1 = scales.next(B, rate) if (training and rate > 0.0) else None
        s2 = scales.next(B, rate) if (training and rate > 0.0) else None
        y, saved = lewin_forward_train(ly, x, s1, s2, gelu_act=gelu_act, qk_scale=qk_scale)
        x = tape.record(y, (lambda ly, saved, xin: (lambda g: [(xin, lewin_backward(ly, saved, g))]))(ly, saved, x))
    return x


def _t_down(tape: Tape, conv, x: torch.Tensor) -> torch.Tensor:
    """``DownsampleLayer`` (``layers/downsample.py:19-30``): conv 4x4 stride 2 pad 1."""
    from . import ops
    N, H, W, _ = x.shape
    out = torch.empty((N, H // 2, W // 2, conv.weight.shape[0]), device=x.device, dtype=x.dtype)
    w = conv.weight.detach().permute(0, 2, 3, 1).reshape(conv.weight.shape[0], -1).to(x.dtype).contiguous()
    ops.conv_gemm([x], w, out, kh=4, kw=4, stride=2, pad=1, bias=conv.bias.detach().contiguous())

    def bwd(g):
        dw, db = ops.conv_wgrad(x, g, 4, 4, 2, 1)
        _accumulate(conv.weight, dw), _accumulate(conv.bias, db)
        return [(x, down4_dgrad(g, conv.weight))]
    return tape.record(out, bwd)


def _t_up(tape: Tape, deconv, x: torch.Tensor) -> torch.Tensor:
    """``UpsampleLayer`` (``layers/upsample.py:19-30``): ConvTranspose 2x2 stride 2 as a per-pixel GEMM with the 2x2 scatter store."""
    from . import ops, _lib as L
    N, H, W, _ = x.shape
    Co = deconv.weight.shape[1]
    out = torch.empty((N, 2 * H, 2 * W, Co), device=x.device, dtype=x.dtype)
    w = deconv.weight.detach().permute(2, 3, 1, 0).reshape(-1, deconv.weight.shape[0]).to(x.dtype).contiguous()   # rows (i,j,co)
    ops.conv_gemm([x], w, out, bias=deconv.bias.detach().repeat(4).contiguous(), store_mode=L.STORE_CONVT2)

    def bwd(g):
        dw, db = convT2_wgrad(x, g)
        _accumulate(deconv.weight, dw), _accumulate(deconv.bias, db)
        return [(x, convT2_dgrad(g, deconv.weight))]
    return tape.record(out, bwd)


def _t_cat(tape: Tape, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Channel concat ``[a | b]`` of two channels-last maps (``models/fba_net.py:282,286``); its backward is the two slices."""
    Ca = a.shape[-1]
    return tape.record(torch.cat([a, b], -1), lambda g: [(a, g[..., :Ca]), (b, g[..., Ca:])])


def _t_conv(tape: Tape, conv, x: torch.Tensor, act: int = 0, alpha: Optional[torch.nn.Parameter] = None) -> torch.Tensor:
    """Stride-1 ``k x k`` convolution (``layers/conv2d.py:12-46``; ``k`` = 1 or 3, padding ``k // 2``) + bias, optionally followed by
    ReLU / PReLU (``alpha``: the scalar slope parameter) as a stand-alone pass that keeps the pre-activation."""
    from . import ops, _lib as L
    Co, Ci, k, _ = conv.weight.shape
    N, H, W, _ = x.shape
    pre = torch.empty((N, H, W, Co), device=x.device, dtype=x.dtype)
    w = conv.weight.detach().permute(0, 2, 3, 1).reshape(Co, -1).to(x.dtype).contiguous()
    ops.conv_gemm([x], w, pre, kh=k, kw=k, pad=k // 2, bias=conv.bias.detach().contiguous())
    if act == L.ACT_NONE:
        y = pre
    elif act == L.ACT_PRELU:
        y = ops.act_forward(pre, act, alpha=alpha.detach().contiguous())
    else:
        y = ops.act_forward(pre, act)

    def bwd(g):
        if act == L.ACT_PRELU:
            g, dalpha = ops.act_backward(pre, g, act, alpha=alpha.detach().contiguous())
            _accumulate(alpha, dalpha)
        elif act != L.ACT_NONE:
            g = ops.act_backward(pre, g, act)
        dw, db = ops.conv_wgrad(x, g, k, k, 1, k // 2)
        _accumulate(conv.weight, dw), _accumulate(conv.bias, db)
        dx = torch.empty_like(x)
        ops.conv_gemm([g], dgrad_weight(conv.weight, g.dtype), dx, kh=k, kw=k, pad=k - 1 - k // 2)
        return [(x, dx)]
    return tape.record(y, bwd)


def _t_add(tape: Tape, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Residual sum ``a + b`` (``fbanet_drop_path_add`` with unit scales); both operands receive the output's gradient."""
    from . import ops
    ones = torch.ones(a.shape[0], device=a.device, dtype=torch.float32)
    return tape.record(ops.drop_path_add(b, ones, skip=a), lambda g: [(a, g), (b, g)])


def _t_resblock(tape: Tape, rb, x: torch.Tensor) -> torch.Tensor:
    """``ResBlock`` (``blocks/residual.py:21-29``): ``x + conv(relu(conv(x)))``."""
    from . import _lib as L
    return _t_add(tape, x, _t_conv(tape, rb.body[2], _t_conv(tape, rb.body[0], x, L.ACT_RELU)))


def _t_faf_gate(tape: Tape, fu, feat: torch.Tensor) -> torch.Tensor:
    """``FAFBlock.compute_guided_aligned_features`` (``blocks/federated_affinity_fusion.py:67-108``): feat ``[B,F,H,W,E]`` -> gated
    features ``[B,H,W,F*E]`` (pixel-major, the K axis of the 1x1 fusion conv).  The scores ``s_f = wsum (*) feat_f`` the backward
    needs for ``sign(s_f - s_0)`` come from the 3x3 implicit GEMM: hi / lo bf16 rows on the tensor cores (bf16), or a 4-row padded
    fp32 GEMM whose first column is the score (fp32)."""
    from . import ops
    B, Fr, H, W, E = feat.shape
    w1 = fu.temporal_attn1.weight
    wsum = w1.detach().double().sum(0).permute(1, 2, 0).reshape(9, E).to(torch.float32).contiguous()
    gate, gated = ops.faf_gate(feat, wsum, want_gate=True, want_gated=True)
    if feat.dtype == torch.bfloat16:
        score = ops.faf_scores(feat, ops.faf_score_weight(wsum, feat.dtype))                     # [B*F,H,W,2]
    else:
        wpad = torch.zeros((4, 9 * E), device=feat.device, dtype=feat.dtype)
        wpad[0] = wsum.reshape(-1)
        s4 = torch.empty((B * Fr, H, W, 4), device=feat.device, dtype=feat.dtype)
        ops.conv_gemm([feat.view(B * Fr, H, W, E)], wpad, s4, kh=3, kw=3, pad=1)
        score = s4[..., 0].float().contiguous()

    def bwd(g):
        dfeat, dwsum = ops.faf_gate_backward(feat, g.reshape(B, H, W, Fr * E), gate, score, wsum)
        _accumulate(w1, ops.faf_weight_grads(dwsum, w1.shape[0]))          # temporal_attn0 and both biases cancel: zero gradient
        return [(feat, dfeat)]
    return tape.record(gated, bwd)


def faf_forward_train(fu, feat: torch.Tensor):
    """Training-mode forward of the whole ``FAFBlock`` (``blocks/federated_affinity_fusion.py:166-182``: gate, 1x1 fusion + PReLU,
    the residual-block hourglass with its 4x4 s2 / transposed-conv resampling, ``fusion_tail`` and the skip) on a :class:`Tape`.
    ``fu``: the model's ``fusion`` module; ``feat`` ``[B,F,H,W,E]`` channels-last.  Returns ``(fused [B,H,W,E], tape)``."""
    tape = Tape()
    return _faf_on_tape(tape, fu, feat), tape


def _faf_on_tape(tape: Tape, fu, feat: torch.Tensor) -> torch.Tensor:
    from . import _lib as L
    gated = _t_faf_gate(tape, fu, feat)
    z = _t_conv(tape, fu.feature_fusion[0], gated, L.ACT_PRELU, fu.feature_fusion[1].weight)
    rb = fu.res_blocks
    r0 = _t_resblock(tape, rb[0][1], _t_resblock(tape, rb[0][0], z))
    r1 = _t_resblock(tape, rb[1][1], _t_resblock(tape, rb[1][0], _t_down(tape, fu.downsample0, r0)))
    r2 = _t_resblock(tape, rb[2][1], _t_resblock(tape, rb[2][0], _t_down(tape, fu.downsample1, r1)))
    r3 = _t_resblock(tape, rb[3][1], _t_resblock(tape, rb[3][0], _t_cat(tape, _t_up(tape, fu.upsample0, r2), r1)))
    r4 = _t_resblock(tape, rb[4][1], _t_resblock(tape, rb[4][0], _t_cat(tape, _t_up(tape, fu.upsample1, r3), r0)))
    return _t_add(tape, _t_conv(tape, fu.fusion_tail, r4), z)


_HG_BLOCKS = ("encoderlayer_0", "encoderlayer_1", "conv", "decoderlayer_0", "decoderlayer_1")


class DropPathDraws:
    """All DropPath multipliers of one training step, drawn in one go: the host generator is consumed in exactly the order the
    forward will ask for them (per hourglass: the five blocks in execution order, per layer the attention branch then the LeFF
    branch; layers with rate 0 draw nothing), stacked ``[calls, batch]`` and sent to the device with ONE pinned asynchronous copy
    instead of ~40 small synchronous ones.  :meth:`next` hands out the rows in that order."""

    def __init__(self, model, hourglasses: int, batch: int, generator: Optional[torch.Generator], device, training: bool = True):
        rates = drop_path_rates(tuple(model.depths), model.drop_path_rate)
        self.rates = [r for _ in range(hourglasses) for b in _HG_BLOCKS for r in rates[b] for _ in (0, 1) if r > 0.0] if training else []
        rows = [drop_path_scales(batch, r, generator) for r in self.rates]
        host = torch.stack(rows) if rows else torch.empty((0, batch), dtype=torch.float32)
        if torch.device(device).type == "cuda" and rows:
            host = host.pin_memory()
        self.scales = host.to(device, non_blocking=True)
        self.k = 0

    def next(self, batch: int, rate: float) -> torch.Tensor:
        assert self.k < len(self.rates) and self.rates[self.k] == rate and self.scales.shape[1] == batch, "DropPath draws out of order"
        self.k += 1
        return self.scales[self.k - 1]


def hourglass_forward_train(model, hg: str, y: torch.Tensor, generator: Optional[torch.Generator] = None, training: bool = True):
    """Training-mode forward of the first LeWin hourglass (``models/fba_net.py:271-287``; ``hg = "HG1"``): encoder blocks, 4x4 s2
    downsamples, bottleneck, 2x2 transposed-conv upsamples concatenated with the encoder outputs, decoder blocks -- ten LeWin layers
    with the per-layer DropPath rates of :func:`drop_path_rates`.  ``y`` ``[B,S,S,E]`` channels-last.  Returns ``(deconv1, tape)``;
    ``tape.backward(deconv1, d_deconv1)`` accumulates every parameter gradient and returns ``{id(y): dy}``."""
    tape = Tape()
    draws = DropPathDraws(model, 1, y.shape[0], generator, y.device, training)
    return _hourglass_on_tape(tape, model, hg, y, None, draws, training)[0], tape


def _hourglass_on_tape(tape: Tape, model, hg: str, y: torch.Tensor, prev, draws: DropPathDraws, training: bool):
    """``prev``: ``(up0, conv1, up1, conv0)`` of the first hourglass -- the second one (``models/fba_net.py:294-310``) feeds its decoders
    ``output_proj_HG2_k(cat[prev pair, own pair])`` (conv3x3 + PReLU) instead of the plain concat."""
    from . import _lib as L
    g = lambda n: getattr(model, n)
    rates = drop_path_rates(tuple(model.depths), model.drop_path_rate)
    act, qks = model.gelu_act, model.qk_scale            # the same GELU flavour and attention scale the inference path uses
    conv0 = _t_block(tape, g(f"{hg}_encoderlayer_0"), y, rates["encoderlayer_0"], draws, training, act, qks)
    pool0 = _t_down(tape, g(f"{hg}_downsample_0").conv[0], conv0)
    conv1 = _t_block(tape, g(f"{hg}_encoderlayer_1"), pool0, rates["encoderlayer_1"], draws, training, act, qks)
    pool1 = _t_down(tape, g(f"{hg}_downsample_1").conv[0], conv1)
    conv2 = _t_block(tape, g(f"conv_{hg}"), pool1, rates["conv"], draws, training, act, qks)
    up0 = _t_up(tape, g(f"{hg}_upsample_0").deconv[0], conv2)
    d0_in = _t_cat(tape, up0, conv1)
    if prev is not None:
        pr = model.output_proj_HG2_0.proj
        d0_in = _t_conv(tape, pr[0], _t_cat(tape, _t_cat(tape, prev[0], prev[1]), d0_in), L.ACT_PRELU, pr[1].weight)
    deconv0 = _t_block(tape, g(f"{hg}_decoderlayer_0"), d0_in, rates["decoderlayer_0"], draws, training, act, qks)
    up1 = _t_up(tape, g(f"{hg}_upsample_1").deconv[0], deconv0)
    d1_in = _t_cat(tape, up1, conv0)
    if prev is not None:
        pr = model.output_proj_HG2_1.proj
        d1_in = _t_conv(tape, pr[0], _t_cat(tape, _t_cat(tape, prev[2], prev[3]), d1_in), L.ACT_PRELU, pr[1].weight)
    deconv1 = _t_block(tape, g(f"{hg}_decoderlayer_1"), d1_in, rates["decoderlayer_1"], draws, training, act, qks)
    return deconv1, (up0, conv1, up1, conv0)


# ------------------------------------------------------------------------------------------------------------------------------
# the whole model in training mode, and one training step (BASELINE config 5; train.py.bak:163-170, train.py:28-66)
# ------------------------------------------------------------------------------------------------------------------------------
def _t_head(tape: Tape, conv, x4: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Head conv (``models/fba_net.py:88,255``) straight from the planar burst: frames ``[N,Cin,S,S]`` fp32 -> channels-last with the
    channels zero-padded to a 16-byte pixel, 3x3 conv with the weight padded alike.  The burst needs no gradient."""
    from . import ops
    Co, Ci = conv.weight.shape[:2]
    cp = 4 if dtype == torch.float32 else 8
    N, _, H, W = x4.shape
    xn = ops.to_nhwc(x4, cp, dtype)
    w = torch.nn.functional.pad(conv.weight.detach().permute(0, 2, 3, 1), (0, cp - Ci)).reshape(Co, -1).to(dtype).contiguous()
    out = torch.empty((N, H, W, Co), device=x4.device, dtype=dtype)
    ops.conv_gemm([xn], w, out, kh=3, kw=3, pad=1, bias=conv.bias.detach().contiguous())

    def bwd(g):
        dw, db = ops.conv_wgrad(xn, g, 3, 3, 1, 1)
        _accumulate(conv.weight, dw[:, :Ci]), _accumulate(conv.bias, db)
        return []
    return tape.record(out, bwd)


def _t_ps_conv(tape: Tape, conv, x: torch.Tensor) -> torch.Tensor:
    """conv3x3 ``E -> 4E`` + ``PixelShuffle(2)`` (``blocks/upsampler.py:22-32``): the shuffle is the ConvT-style scatter store of the
    GEMM whose rows are re-ordered from ``4c+2i+j`` to ``(2i+j)*E + c``."""
    from . import ops, _lib as L
    C4, Ci = conv.weight.shape[:2]
    N, H, W, _ = x.shape
    w = conv.weight.detach().permute(0, 2, 3, 1).reshape(C4 // 4, 4, -1).permute(1, 0, 2).reshape(C4, -1).to(x.dtype).contiguous()
    b = conv.bias.detach().reshape(C4 // 4, 4).t().reshape(-1).contiguous()
    out = torch.empty((N, 2 * H, 2 * W, C4 // 4), device=x.device, dtype=x.dtype)
    ops.conv_gemm([x], w, out, kh=3, kw=3, pad=1, bias=b, store_mode=L.STORE_CONVT2)

    def bwd(g):
        dw, db = pixel_shuffle_conv_wgrad(x, g)
        _accumulate(conv.weight, dw), _accumulate(conv.bias, db)
        return [(x, pixel_shuffle_conv_dgrad(g, conv.weight))]
    return tape.record(out, bwd)


def _t_final(tape: Tape, conv, x: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """Last conv ``E -> C_in`` + the bilinear x4 base frame (``models/fba_net.py:315-320``): planar fp32 ``[B,C_in,4S,4S]`` out.  The
    incoming gradient is planar fp32 (``fbanet_train_loss_sm100``); it is re-laid channels-last (padded to a 16-byte pixel) for the
    weight-gradient kernel and for the flipped-weight data gradient.  The base path has no parameters."""
    from . import ops, _lib as L
    Co, Ci = conv.weight.shape[:2]
    N, H, W, _ = x.shape
    cp = 4 if x.dtype == torch.float32 else 8
    w = conv.weight.detach().permute(0, 2, 3, 1).reshape(Co, -1).to(x.dtype).contiguous()
    out = torch.empty((N, Co, H, W), device=x.device, dtype=torch.float32)
    ops.conv_gemm([x], w, out, kh=3, kw=3, pad=1, bias=conv.bias.detach().contiguous(), store_mode=L.STORE_NCHW_BASE, base=base, cout_store=Co)

    def bwd(g):
        gn = ops.to_nhwc(g.contiguous(), cp, x.dtype)                                            # [N,H,W,cp], channels >= Co are zero
        dw, db = ops.conv_wgrad(x, gn, 3, 3, 1, 1)
        _accumulate(conv.weight, dw[:Co]), _accumulate(conv.bias, db[:Co])
        wp = torch.nn.functional.pad(conv.weight.detach(), (0, 0, 0, 0, 0, 0, 0, cp - Co))       # zero rows for the padded channels
        dx = torch.empty_like(x)
        ops.conv_gemm([gn], dgrad_weight(wp, x.dtype), dx, kh=3, kw=3, pad=1)
        return [(x, dx)]
    return tape.record(out, bwd)


def model_forward_train(model, burst: torch.Tensor, generator: Optional[torch.Generator] = None, training: bool = True):
    """The whole ``FBANetModel.__call__`` (``models/fba_net.py:242-322``) in training mode on a :class:`Tape`: ``burst`` fp32
    ``[B,F,C_in,S,S]`` on the GPU -> ``(restored [B,C_in,4S,4S] fp32, tape)``.  Every arithmetic step is a C-ABI op; activations
    live in the model's compute dtype.  ``tape.backward(restored, d_restored)`` accumulates the gradient of every parameter the
    output depends on (``fusion.temporal_attn0`` and the two embedding biases cancel out of the gate as written and stay untouched)."""
    from . import _lib as L
    B, Fr, Cin, S, _ = burst.shape
    E, dt = model.embed_dim, model.compute_dtype
    tape = Tape()
    burst = burst.contiguous().float()
    f = _t_head(tape, model.head, burst.view(B * Fr, Cin, S, S), dt)
    f = _t_resblock(tape, model.body[1], _t_resblock(tape, model.body[0], f))
    feat = tape.record(f.view(B, Fr, S, S, E), (lambda f: (lambda g: [(f, g.reshape(f.shape))]))(f))
    fused = _faf_on_tape(tape, model.fusion, feat)
    proj = lambda m, x: _t_conv(tape, m.proj[0], x, L.ACT_PRELU, m.proj[1].weight)
    y = proj(model.input_proj, fused)
    draws = DropPathDraws(model, 2, B, generator, burst.device, training)
    d1, prev = _hourglass_on_tape(tape, model, "HG1", y, None, draws, training)
    y1 = proj(model.output_proj, d1)
    d2, _ = _hourglass_on_tape(tape, model, "HG2", y1, prev, draws, training)
    y2 = proj(model.output_proj_2, d2)
    t2 = _t_ps_conv(tape, model.tail[0][2], _t_ps_conv(tape, model.tail[0][0], y2))
    return _t_final(tape, model.tail[1], t2, burst[:, 0]), tape


def train_step(model, flat: FlatParams, burst: torch.Tensor, target: torch.Tensor, lr: float, weight_decay: float = 0.02,
               generator: Optional[torch.Generator] = None, group=None, bucket_bytes: int = 16 << 20):
    """One data-parallel training step of the reference trainer (``train.py.bak:163-170``: forward, ``CharbonnierLoss + 3 GWLoss``,
    backward, ``AdamW`` step; ``DataParallel`` -> one process per GPU + the gradient all-reduce, issued in buckets that overlap the backward).  ``flat = FlatParams(model.parameters())``
    (parameters set to ``requires_grad``).  Returns the loss triple ``(total, charbonnier, gw)`` as a float64 device tensor."""
    from . import ops
    global _REDUCER
    flat.zero_grad()
    restored, tape = model_forward_train(model, burst, generator, training=True)
    loss, d_restored = ops.training_loss(restored, target, clamp_restored=True)      # train.py.bak:167: clamp(restored, 0, 1) first
    flat.begin_reduce(group, bucket_bytes)
    _REDUCER = flat
    try:
        tape.backward(restored, d_restored)          # buckets of finished gradients are all-reduced while this still runs
    finally:
        _REDUCER = None
    scale = flat.finish_reduce()                     # the rest (incl. parameters the loss does not reach), then wait
    flat.adam_step(lr, weight_decay=weight_decay, decoupled=True, grad_scale=scale)
    return loss


def fit(model, flat: FlatParams, batches, nepoch: int, start_epoch: int = 1, lr_initial: float = 1e-4, warmup: bool = True,
        warmup_epochs: int = 3, weight_decay: float = 0.02, model_dir: Optional[str] = None, checkpoint_every: int = 50,
        generator: Optional[torch.Generator] = None, group=None, log=None) -> List[float]:
    """The epoch loop of the reference trainer (``train.py.bak:150-246``) around :func:`train_step`: for every epoch the learning rate
    of ``--warmup`` + cosine (:func:`warmup_cosine_lr`) or ``StepLR`` (:func:`step_lr`), one step per ``(burst, target)`` pair of
    ``batches`` (any re-iterable of device tensors), the summed epoch loss, and -- on rank 0, when ``model_dir`` is given -- the
    reference's checkpoints ``model_latest.pth`` every epoch and ``model_epoch_<e>.pth`` every ``checkpoint_every`` epochs, each
    ``{"epoch", "state_dict", "optimizer"}`` (``:236-245``; resumable through ``utils.load_checkpoint`` / ``load_start_epoch`` /
    ``load_optim``).  Validation, best-PSNR tracking and logging sinks are the caller's (the reference's CLI is out of scope).
    ``generator`` (DropPath): default = a per-rank ``Generator`` seeded ``1234 + rank``.  Under data parallelism every rank must
    supply the SAME number of batches per epoch (asserted when ``batches`` has a length).  Returns the list of epoch losses."""
    import os
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    rank0 = rank == 0
    if generator is None:
        # the reference seeds every process alike (train.py.bak:56-59, 1234); data-parallel ranks hold DIFFERENT bursts, so each
        # draws its own DropPath pattern: seed + rank
        generator = torch.Generator().manual_seed(1234 + rank)
    if distributed and dist.get_world_size(group) > 1 and hasattr(batches, "__len__"):
        # every rank must take the same number of steps: a rank that runs out early would leave the others in finish_reduce
        n = torch.tensor([len(batches)], dtype=torch.int64, device=flat.data.device)
        lo, hi = n.clone(), n.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        assert lo.item() == hi.item(), f"ranks hold different numbers of batches ({lo.item()}..{hi.item()}): pad or drop the last ones"
    history = []
    for epoch in range(start_epoch, nepoch + 1):
        lr = warmup_cosine_lr(epoch, lr_initial, nepoch, warmup_epochs) if warmup else step_lr(epoch, lr_initial)
        total = None
        for burst, target in batches:
            loss = train_step(model, flat, burst, target, lr, weight_decay, generator, group)
            total = loss[0].clone() if total is None else total + loss[0]          # stays on the device: no sync per step
        history.append(float(total) if total is not None else 0.0)
        if log is not None:
            log(epoch, lr, history[-1])
        if model_dir is not None and rank0:
            state = {"epoch": epoch, "state_dict": model.state_dict(), "optimizer": flat.state_dict()}
            torch.save(state, os.path.join(model_dir, "model_latest.pth"))
            if epoch % checkpoint_every == 0:
                torch.save(state, os.path.join(model_dir, "model_epoch_{}.pth".format(epoch)))
    return history


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
