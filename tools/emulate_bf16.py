"""CPU emulation of the bf16 path's rounding points on the oracle (no GPU): where does the PSNR delta of the parity test come from?

The oracle (fp32, torch CPU) is run with (a) conv / linear weights rounded to bf16, (b) the output of every conv / linear / LayerNorm
rounded to bf16 (the kernels store bf16 activations), or both, per stage group, and the test's metric
|PSNR(variant, gt) - PSNR(fp32, gt)| is printed per seed.  `--fold` emulates the folded LayerNorm (no rounding of the LayerNorm
output, consumer weights replaced by the centred, gamma-scaled bf16 weights of ops.fold_layernorm).

    python tools/emulate_bf16.py --seeds 0 1 2 3 4 5 6 7 --variants all w a fold
"""
import argparse
import math
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.fbanet_oracle import build_oracle, psnr  # noqa: E402
from oracle import fbanet_oracle as O  # noqa: E402

FULL = dict(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10)


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def hilo(t):
    """what a hi + lo bf16 pair represents"""
    hi = bf(t)
    return hi + bf(t - hi)


def bf_rowsum(w):
    """bf16 rounding that keeps every output channel's weight SUM (the response to the common mode of its inputs): the row's
    rounding residual is pushed into its smallest-magnitude elements, whose ulps are finest (as ops.fold_layernorm does)."""
    shp = w.shape
    w2 = w.reshape(shp[0], -1).double()
    wf = w2.to(torch.bfloat16)
    for t in (8, 1):
        r = wf.double().sum(1, keepdim=True) - w2.sum(1, keepdim=True)
        idx = wf.abs().float().topk(min(t, wf.shape[1]), dim=1, largest=False).indices
        wf.scatter_(1, idx, (wf.gather(1, idx).double() - r / idx.shape[1]).to(torch.bfloat16))
    return wf.float().reshape(shp)


ROUND = bf


def burst(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, cfg["num_frames"], cfg["in_channels"], cfg["img_size"], cfg["img_size"], generator=g)


def fold_ln(w, b, gamma, beta, exact=False):
    """ops.fold_layernorm restated (bf16); ``exact``: centred weights kept in fp32 (a hi + lo bf16 pair carries 16 mantissa bits)."""
    w64, g64, be64 = w.double(), gamma.double(), beta.double()
    wg = w64 * g64[None, :]
    if exact:
        return hilo((wg - wg.mean(1, keepdim=True)).float()), (w64 @ be64 + b.double()).float()
    wf = (wg - wg.mean(1, keepdim=True)).to(torch.bfloat16)
    for t in (8, 1):
        r = wf.double().sum(1, keepdim=True)
        idx = wf.abs().float().topk(min(t, wf.shape[1]), dim=1, largest=False).indices
        wf.scatter_(1, idx, (wf.gather(1, idx).double() - r / idx.shape[1]).to(torch.bfloat16))
    return wf.float(), (w64 @ be64 + b.double()).float()


class FoldedLN(nn.Module):
    """LayerNorm without affine and without output rounding: (x - mean) * rstd; the consumer carries gamma / beta."""

    def __init__(self, ln):
        super().__init__()
        self.eps, self.shape = ln.eps, ln.normalized_shape

    def forward(self, x):
        return torch.nn.functional.layer_norm(x, self.shape, None, None, self.eps)


def apply_variant(m, weights=None, acts=None, fold=False, exact=()):
    """weights / acts: predicate(name) -> bool selecting the leaf modules whose weights / outputs are rounded."""
    hooks = []
    if fold:
        for name, mod in list(m.named_modules()):
            if isinstance(mod, O.LeWinLayer):
                # norm1 -> attn.qkv (to_q, to_kv), norm2 -> mlp.linear1
                att, mlp = mod.attn, mod.mlp
                ex = any(p in name for p in exact)
                for lin in (att.qkv.to_q, att.qkv.to_kv):
                    wf, b = fold_ln(lin.weight.data, lin.bias.data, mod.norm1.weight.data, mod.norm1.bias.data, ex)
                    lin.weight.data, lin.bias.data = wf, b
                    lin._folded = True
                l1 = mlp.linear1[0] if isinstance(mlp.linear1, nn.Sequential) else mlp.linear1
                wf, b = fold_ln(l1.weight.data, l1.bias.data, mod.norm2.weight.data, mod.norm2.bias.data, ex)
                l1.weight.data, l1.bias.data = wf, b
                l1._folded = True
                mod.norm1, mod.norm2 = FoldedLN(mod.norm1), FoldedLN(mod.norm2)
    for name, mod in m.named_modules():
        leaf = isinstance(mod, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d, nn.LayerNorm))
        if not leaf:
            continue
        if weights and weights(name) and not isinstance(mod, nn.LayerNorm) and not getattr(mod, "_folded", False):
            if "dwconv" in name:
                continue                                   # depthwise taps stay fp32 in the kernels
            mod.weight.data = hilo(mod.weight.data) if (any(p in name for p in exact) or name == "tail.1") else ROUND(mod.weight.data)
        if acts and acts(name):
            hooks.append(mod.register_forward_hook(lambda _m, _i, out: bf(out)))
    return hooks


def run(seed, variant, exact=()):
    o = build_oracle(seed, **FULL)
    x = burst(FULL, 1, seed)
    with torch.no_grad():
        ref = o(x)
    g = torch.Generator().manual_seed(7)
    gt = (torch.nn.functional.interpolate(x[:, 0], scale_factor=4, mode="bilinear", align_corners=False)
          + 0.05 * torch.randn(ref.shape, generator=g)).clamp(0, 1)
    res = {}
    for v in variant:
        m = build_oracle(seed, **FULL)
        every = lambda n: True
        if v == "all":
            apply_variant(m, every, every)
        elif v == "allsum":
            global ROUND
            ROUND = bf_rowsum
            apply_variant(m, every, every)
            ROUND = bf
        elif v == "foldsum":
            ROUND = bf_rowsum
            apply_variant(m, every, every, fold=True)
            ROUND = bf
        elif v == "allx":
            apply_variant(m, every, every, exact=exact)
        elif v == "foldx":
            apply_variant(m, every, every, fold=True, exact=exact)
        elif v == "w":
            apply_variant(m, every, None)
        elif v == "a":
            apply_variant(m, None, every)
        elif v == "fold":
            apply_variant(m, every, every, fold=True)
        elif v.startswith("w:"):      # weights of the modules whose name contains the pattern, activations everywhere
            pat = v[2:]
            apply_variant(m, lambda n: pat in n, None)
        elif v.startswith("a:"):
            pat = v[2:]
            apply_variant(m, None, lambda n: pat in n)
        elif v.startswith("not:"):    # everything rounded except modules whose name contains the pattern
            pat = v[4:]
            apply_variant(m, lambda n: pat not in n, lambda n: pat not in n)
        else:
            raise SystemExit(v)
        with torch.no_grad():
            got = m(x)
        s = psnr(got.clamp(0, 1), gt) - psnr(ref.clamp(0, 1), gt)
        res[v] = (s, psnr(got, ref))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--variants", nargs="+", default=["all", "w", "a", "fold"])
    ap.add_argument("--exact", nargs="*", default=[], help="allx / foldx: modules (name substrings) whose weights are hi + lo bf16 pairs")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    for s in a.seeds:
        r = run(s, a.variants, tuple(a.exact))
        print("seed", s, "  ".join(f"{k}: {d:+.4f} dB ({p:.1f})" for k, (d, p) in r.items()), flush=True)


if __name__ == "__main__":
    main()
