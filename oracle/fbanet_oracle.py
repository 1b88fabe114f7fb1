"""CPU oracle for the FBANet ``BaseModel`` burst-SR forward path.  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch (CPU, fp32) restatement of the reference forward
``/root/reference/fba_net/models/fba_net.py:242-322`` and everything it calls.  It is *not* part of
the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package (``fbanet_b200``) never does, and it
fails loudly when its CUDA library is missing.

PARITY: PINNED LAYER BY LAYER, UNPINNED AS A WHOLE.  The reference ships no tests, no golden vectors and no checkpoint, JAX
is not installed in the build image, and the reference's whole forward cannot execute as written.
What CAN execute of it is executed: `tests/golden/make_golden_layers.py` / `make_golden_model.py` import the unmodified
`fba_net.layers.*`, `fba_net.blocks.*` and `fba_net.models.fba_net` from /root/reference on top of numpy stand-ins for the ~20
jax / equinox primitives they call (`tests/golden/jaxshim/`) and record the outputs of every layer that runs as written --
convolutions (3x3, 1x1, 4x4 s2, depthwise, transposed), ResBlock, InputProjLayer, LinearProjectionLayer (heads 1/2/4),
WindowAttentionLayer (heads 1), the whole FAFBlock, FBANetLayer's mask / cyclic shift / window partition / reverse, the unshifted
FBANetLayer, the model's constructor (structure, parameter counts) and its `__call__` wiring over stub sub-modules.  The classes and
functions below reproduce those vectors (`tests/test_oracle_reference_layers.py`; fixtures `tests/golden/layers_reference.npz`,
`model_structure_reference.json`, `model_wiring_reference.npz`).  Still pinned only by reading + closed forms: the places where the
reference does not execute (Appendix A-1 mask add, A-2 index, A-3 heads > 1, A-4 residuals, A-5/6 LeFF reshape, A-17/18/19 x4 tail,
PixelShuffle order, bilinear base) and the primitives' semantics the stand-ins restate from the Equinox / JAX documentation.
Further independent pins: numpy window index / shift-mask formulas, ``cv2.warpPerspective`` in 1/32-px mode, float64
restatements, scipy / cv2 for the alignment front ends, and the reference's own torch-only tiling / loss / checkpoint code.

Every place where the reference is not executable follows the decision register of SURVEY.md
Appendix A ("A-n" below).  Layout: channels-first ``[N,C,H,W]`` images and ``[N,T,C]`` tokens
internally (token ``t = y*W + x``), i.e. the convention of the reference's torch-side callers
(``test_in_any_resolution.py:70-93``); the reference's JAX code is channels-last, which is the same
maths (A-11).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------------
def gelu_fn(kind: str):
    """A-13: ``jax.nn.gelu`` defaults to the tanh approximation (blocks/fba_net.py:26)."""
    if kind == "tanh":
        return lambda u: F.gelu(u, approximate="tanh")
    if kind == "erf":
        return lambda u: F.gelu(u)
    raise ValueError(kind)


class ResBlock(nn.Module):
    """blocks/residual.py:21-29 -- ``x + conv(relu(conv(x)))``."""

    def __init__(self, c: int):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1), nn.ReLU(), nn.Conv2d(c, c, 3, 1, 1))

    def forward(self, x):
        return self.body(x) + x


def relative_position_index(w: int) -> torch.Tensor:
    """A-2: Swin index ``(dy + w-1)*(2w-1) + (dx + w-1)`` (window_attention.py:70-90 as intended)."""
    coords = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(w), indexing="ij"))  # 2,w,w
    flat = coords.flatten(1)  # 2, w*w
    rel = flat[:, :, None] - flat[:, None, :]  # 2, N, N   (i minus j)
    rel = rel.permute(1, 2, 0).contiguous()
    rel[:, :, 0] += w - 1
    rel[:, :, 1] += w - 1
    rel[:, :, 0] *= 2 * w - 1
    return rel.sum(-1)  # N,N long


def shift_attn_mask(H: int, W: int, win: int, shift: int) -> torch.Tensor:
    """layers/fba_net.py:149-184 -- 9-region id image, partitioned, ``-100`` where ids differ."""
    img = torch.zeros(H, W)
    cnt = 0
    for hs in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
        for ws in (slice(0, -win), slice(-win, -shift), slice(-shift, None)):
            img[hs, ws] = cnt
            cnt += 1
    m = window_partition(img[None, :, :, None], win).squeeze(-1)  # nW, N
    diff = m[:, None, :] - m[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))  # nW,N,N


def window_partition(x: torch.Tensor, win: int) -> torch.Tensor:
    """layers/fba_net.py:113-124 -- ``[B,H,W,C] -> [B*nW, win*win, C]``; window id row-major."""
    B, H, W, C = x.shape
    x = x.view(B, H // win, win, W // win, win, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, win * win, C)


def window_reverse(w: torch.Tensor, win: int, B: int, H: int, W: int) -> torch.Tensor:
    """layers/fba_net.py:126-137."""
    C = w.shape[-1]
    x = w.view(B, H // win, W // win, win, win, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


class LinearProjection(nn.Module):
    """layers/linear_projection.py:24-44 -- ``to_q`` [d,d], ``to_kv`` [2d,d]; k rows first, then v."""

    def __init__(self, dim: int, heads: int, bias: bool = True):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=bias)
        self.to_kv = nn.Linear(dim, 2 * dim, bias=bias)

    def forward(self, x):  # x: [Bw, N, d]
        Bw, N, d = x.shape
        h = self.heads
        q = self.to_q(x).view(Bw, N, h, d // h).permute(0, 2, 1, 3)
        kv = self.to_kv(x).view(Bw, N, 2, h, d // h).permute(2, 0, 3, 1, 4)
        return q, kv[0], kv[1]


class WindowAttention(nn.Module):
    """layers/window_attention.py:140-248 with A-1, A-2, A-3, A-22."""

    def __init__(self, dim: int, win: int, heads: int, qk_scale=None):
        super().__init__()
        self.dim, self.win, self.heads = dim, win, heads
        self.scale = qk_scale or (dim // heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * win - 1) ** 2, heads))
        self.register_buffer("relative_position_index", relative_position_index(win))
        self.qkv = LinearProjection(dim, heads)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, mask=None):  # x [Bw,N,d]; mask [nW,N,N] or None
        Bw, N, d = x.shape
        q, k, v = self.qkv(x)
        q = q * self.scale  # window_attention.py:174 (after the bias add of to_q)
        attn = q @ k.transpose(-2, -1)  # Bw,h,N,N
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)]
        bias = bias.view(N, N, self.heads).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if mask is not None:  # A-1
            nW = mask.shape[0]
            attn = attn.view(Bw // nW, nW, self.heads, N, N) + mask[None, :, None]
            attn = attn.view(Bw, self.heads, N, N)
        attn = torch.softmax(attn, dim=-1)
        out = (attn @ v).transpose(1, 2).reshape(Bw, N, d)
        return self.proj(out)


class LeFF(nn.Module):
    """layers/locally_enhanced_feed_forward.py:25-60 with A-5/A-6 (whole H x W map, depthwise 3x3)."""

    def __init__(self, dim: int, hidden: int, gelu: str):
        super().__init__()
        self.act = gelu_fn(gelu)
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden))
        self.dwconv = nn.Sequential(nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden))
        self.linear2 = nn.Sequential(nn.Linear(hidden, dim))

    def forward(self, x, H, W):  # x [B,T,d]
        B, T, d = x.shape
        u = self.act(self.linear1(x))
        u = u.transpose(1, 2).reshape(B, -1, H, W)
        u = self.act(self.dwconv(u))
        u = u.flatten(2).transpose(1, 2)
        return self.linear2(u)


class LeWinLayer(nn.Module):
    """layers/fba_net.py:50-250 (``FBANetLayer``) with A-4 residuals and A-23 window clamp."""

    def __init__(self, dim, res, heads, win, shift, mlp_ratio, gelu):
        super().__init__()
        H, W = res
        if min(H, W) <= win:  # layers/fba_net.py:55-66
            shift, win = 0, min(H, W)
        assert H % win == 0 and W % win == 0
        self.dim, self.res, self.win, self.shift = dim, res, win, shift
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, win, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = LeFF(dim, int(dim * mlp_ratio), gelu)

    def forward(self, x):  # [B,T,d]
        H, W = self.res
        B, T, d = x.shape
        mask = shift_attn_mask(H, W, self.win, self.shift) if self.shift > 0 else None
        skip = x
        y = self.norm1(x).view(B, H, W, d)
        if self.shift > 0:
            y = torch.roll(y, shifts=(-self.shift, -self.shift), dims=(1, 2))
        yw = window_partition(y, self.win)
        aw = self.attn(yw, mask)
        y = window_reverse(aw, self.win, B, H, W)
        if self.shift > 0:
            y = torch.roll(y, shifts=(self.shift, self.shift), dims=(1, 2))
        x = skip + y.view(B, T, d)  # A-4
        x = x + self.mlp(self.norm2(x), H, W)
        return x


class LeWinBlock(nn.Module):
    """blocks/fba_net.py:35-65 -- ``depth`` layers, shift 0 / win//2 alternating."""

    def __init__(self, dim, res, depth, heads, win, mlp_ratio, gelu):
        super().__init__()
        self.blocks = nn.ModuleList(
            [LeWinLayer(dim, res, heads, win, 0 if i % 2 == 0 else win // 2, mlp_ratio, gelu) for i in range(depth)]
        )

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class Downsample(nn.Module):
    """layers/downsample.py:19-30 + downsample_flatten.py:6-13 -- tokens -> conv4x4 s2 p1 -> tokens."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 4, 2, 1))

    def forward(self, x, H, W):
        B, T, C = x.shape
        y = self.conv(x.transpose(1, 2).reshape(B, C, H, W))
        return y.flatten(2).transpose(1, 2)


class Upsample(nn.Module):
    """layers/upsample.py:19-30 + upsample_flatten.py:6-12 -- tokens -> convT 2x2 s2 -> tokens."""

    def __init__(self, cin, cout):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose2d(cin, cout, 2, 2))

    def forward(self, x, H, W):
        B, T, C = x.shape
        y = self.deconv(x.transpose(1, 2).reshape(B, C, H, W))
        return y.flatten(2).transpose(1, 2)


class Proj(nn.Module):
    """input_projection.py:30-43 / output_projection(_hwc).py -- conv3x3 + scalar PReLU (A-14)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1), nn.PReLU())

    def forward(self, x):  # image in, image out
        return self.proj(x)


class FAFBlock(nn.Module):
    """blocks/federated_affinity_fusion.py:18-182 (A-12: affinity = sum over channels of a difference)."""

    def __init__(self, nf: int, frames: int):
        super().__init__()
        self.nf, self.frames = nf, frames
        self.temporal_attn0 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.temporal_attn1 = nn.Conv2d(nf, nf, 3, 1, 1)
        self.feature_fusion = nn.Sequential(nn.Conv2d(nf * frames, nf, 1, 1, 0), nn.PReLU(init=0.1))
        self.downsample0 = nn.Conv2d(nf, 2 * nf, 4, 2, 1)
        self.downsample1 = nn.Conv2d(2 * nf, 4 * nf, 4, 2, 1)
        self.upsample0 = nn.ConvTranspose2d(4 * nf, 2 * nf, 2, 2)
        self.upsample1 = nn.ConvTranspose2d(4 * nf, nf, 2, 2)
        self.res_blocks = nn.ModuleList(
            [nn.Sequential(ResBlock(nf * m), ResBlock(nf * m)) for m in (1, 2, 4, 4, 2)]
        )
        self.fusion_tail = nn.Conv2d(2 * nf, nf, 3, 1, 1)

    def guided(self, feat):  # feat [B,F,C,H,W]  (:67-108)
        B, Fr, C, H, W = feat.shape
        emb_ref = self.temporal_attn0(feat[:, 0])  # :79
        emb = self.temporal_attn1(feat.reshape(B * Fr, C, H, W)).view(B, Fr, C, H, W)  # :81
        aff = (emb - emb_ref[:, None]).sum(2)  # :84-88  [B,F,H,W]
        diffs = (aff[:, 1:] - aff[:, :1]).abs()  # :91
        gate = torch.sigmoid(diffs)[:, :, None]  # :95-99
        return torch.cat([feat[:, :1], feat[:, 1:] * gate], 1), gate.squeeze(2)  # :102-105

    def fuse(self, g):  # :110-164
        B, Fr, C, H, W = g.shape
        z = self.feature_fusion(g.reshape(B, Fr * C, H, W))  # channel = f*C + c  (:121-128)
        r0 = self.res_blocks[0](z)
        r1 = self.res_blocks[1](self.downsample0(r0))
        r2 = self.res_blocks[2](self.downsample1(r1))
        r3 = self.res_blocks[3](torch.cat([self.upsample0(r2), r1], 1))  # [up, skip]
        r4 = self.res_blocks[4](torch.cat([self.upsample1(r3), r0], 1))
        return self.fusion_tail(r4) + z, z

    def forward(self, feat):
        g, _ = self.guided(feat)
        out, _ = self.fuse(g)
        return out


# ----------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------
class OracleBaseModel(nn.Module):
    """models/fba_net.py:30-322 (``FBANetModel``) as a torch CPU module.

    ``forward(burst[B,T,C,H,W]) -> [B,C,4H,4W]`` -- the signature the reference's torch callers use
    (test_in_any_resolution.py:85).  A-9 (``in_channels`` generalised), A-17 (x4 tail), A-18 (torch
    PixelShuffle order), A-19 (bilinear x4 base, half-pixel centres), A-24 (HG2 reuses heads[0,1,4..6]).
    """

    def __init__(
        self,
        num_frames=14,
        img_size=160,
        in_channels=3,
        embed_dim=64,
        depths=(2, 2, 2, 2, 2, 2, 2, 2, 2),
        heads=(1, 2, 4, 8, 16, 16, 8, 4, 2),
        window_length=10,
        mlp_ratio=4.0,
        gelu="tanh",
    ):
        super().__init__()
        E, S, w = embed_dim, img_size, window_length
        self.num_frames, self.img_size, self.in_channels, self.embed_dim = num_frames, img_size, in_channels, E
        self.head = nn.Conv2d(in_channels, E, 3, 1, 1)
        self.body = nn.Sequential(ResBlock(E), ResBlock(E))
        self.fusion = FAFBlock(E, num_frames)
        self.input_proj = Proj(E, E)
        self.output_proj = Proj(2 * E, E)
        self.output_proj_2 = Proj(2 * E, E)
        self.output_proj_HG2_0 = Proj(8 * E, 4 * E)
        self.output_proj_HG2_1 = Proj(4 * E, 2 * E)
        kw = dict(win=w, mlp_ratio=mlp_ratio, gelu=gelu)
        for hg in ("HG1", "HG2"):
            setattr(self, f"{hg}_encoderlayer_0", LeWinBlock(E, (S, S), depths[0], heads[0], **kw))
            setattr(self, f"{hg}_downsample_0", Downsample(E, 2 * E))
            setattr(self, f"{hg}_encoderlayer_1", LeWinBlock(2 * E, (S // 2, S // 2), depths[1], heads[1], **kw))
            setattr(self, f"{hg}_downsample_1", Downsample(2 * E, 4 * E))
            setattr(self, f"conv_{hg}", LeWinBlock(4 * E, (S // 4, S // 4), depths[4], heads[4], **kw))
            setattr(self, f"{hg}_upsample_0", Upsample(4 * E, 2 * E))
            setattr(self, f"{hg}_decoderlayer_0", LeWinBlock(4 * E, (S // 2, S // 2), depths[5], heads[5], **kw))
            setattr(self, f"{hg}_upsample_1", Upsample(4 * E, E))
            setattr(self, f"{hg}_decoderlayer_1", LeWinBlock(2 * E, (S, S), depths[6], heads[6], **kw))
        self.tail = nn.Sequential(  # A-17
            nn.Sequential(nn.Conv2d(E, 4 * E, 3, 1, 1), nn.PixelShuffle(2), nn.Conv2d(E, 4 * E, 3, 1, 1), nn.PixelShuffle(2)),
            nn.Conv2d(E, in_channels, 3, 1, 1),
        )

    # -- helpers -----------------------------------------------------------------------------
    @staticmethod
    def _img(tok, S):
        B, T, C = tok.shape
        return tok.transpose(1, 2).reshape(B, C, S, S)

    @staticmethod
    def _tok(img):
        return img.flatten(2).transpose(1, 2)

    def _hourglass(self, hg, y, prev=None, st=None):
        """models/fba_net.py:271-287 (HG1) / :294-310 (HG2).  ``prev`` = (up0, conv1, up1, conv0) of HG1."""
        S = self.img_size
        g = lambda n: getattr(self, n)
        conv0 = g(f"{hg}_encoderlayer_0")(y)
        pool0 = g(f"{hg}_downsample_0")(conv0, S, S)
        conv1 = g(f"{hg}_encoderlayer_1")(pool0)
        pool1 = g(f"{hg}_downsample_1")(conv1, S // 2, S // 2)
        conv2 = g(f"conv_{hg}")(pool1)
        up0 = g(f"{hg}_upsample_0")(conv2, S // 4, S // 4)
        if prev is None:
            d0_in = torch.cat([up0, conv1], -1)  # [up, skip]  (:282)
        else:
            cat = torch.cat([prev[0], prev[1], up0, conv1], -1)  # (:305)
            d0_in = self._tok(self.output_proj_HG2_0(self._img(cat, S // 2)))
        deconv0 = g(f"{hg}_decoderlayer_0")(d0_in)
        up1 = g(f"{hg}_upsample_1")(deconv0, S // 2, S // 2)
        if prev is None:
            d1_in = torch.cat([up1, conv0], -1)  # (:286)
        else:
            cat = torch.cat([prev[2], prev[3], up1, conv0], -1)  # (:309)
            d1_in = self._tok(self.output_proj_HG2_1(self._img(cat, S)))
        deconv1 = g(f"{hg}_decoderlayer_1")(d1_in)
        if st is not None:
            for k, v in dict(conv0=conv0, pool0=pool0, conv1=conv1, pool1=pool1, conv2=conv2, up0=up0,
                             deconv0_in=d0_in, deconv0=deconv0, up1=up1, deconv1_in=d1_in, deconv1=deconv1).items():
                st[f"{hg}.{k}"] = v
        return deconv1, (up0, conv1, up1, conv0)

    def forward_stages(self, x: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        """Forward returning every named intermediate (images ``[B,C,H,W]``, tokens ``[B,T,C]``)."""
        B, Fr, C, H, W = x.shape
        S, E = self.img_size, self.embed_dim
        assert (Fr, C, H, W) == (self.num_frames, self.in_channels, S, S), "input must be [B,num_frames,C_in,S,S]"  # :244 / A-9
        st: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        x_base = x[:, 0]  # :246
        f = self.head(x.reshape(B * Fr, C, H, W))  # :255
        st["head"] = f.view(B, Fr, E, H, W)
        f = self.body(f)  # :258
        st["body"] = f.view(B, Fr, E, H, W)
        g, gate = self.fusion.guided(st["body"])
        st["faf.gate"] = gate
        fused, z = self.fusion.fuse(g)
        st["faf.fuse1x1"] = z
        st["fusion"] = fused  # :262
        y = self._tok(self.input_proj(fused))  # :266 (Dropout(0) identity, A-8)
        st["input_proj"] = y
        d1, prev = self._hourglass("HG1", y, None, st)
        y1 = self._tok(self.output_proj(self._img(d1, S)))  # :290
        st["output_proj"] = y1
        d1_2, _ = self._hourglass("HG2", y1, prev, st)
        y2 = self.output_proj_2(self._img(d1_2, S))  # :313
        st["output_proj_2"] = y2
        out2 = self.tail(y2)  # :315
        st["tail"] = out2
        base = F.interpolate(x_base, scale_factor=4, mode="bilinear", align_corners=False)  # :317-318 / A-19
        st["base"] = base
        st["out"] = out2 + base  # :320
        return st

    def forward(self, x):
        return self.forward_stages(x)["out"]


# ----------------------------------------------------------------------------------------------
# deterministic init with the reference's distributions (SURVEY.md 8c "Random-init parity")
# ----------------------------------------------------------------------------------------------
def init_parameters_(model: nn.Module, seed: int = 0) -> nn.Module:
    """conv/linear weight & bias ``U(+-1/sqrt(fan_in))`` (Equinox default), LayerNorm 1/0, PReLU 0.25
    (FAF fusion 0.1, federated_affinity_fusion.py:47), rel-pos table trunc-normal(std .02, +-2 sigma)
    (window_attention.py:55-65,143-145).  Drawn in ``named_parameters()`` order from one torch generator."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            owner = model.get_submodule(name.rsplit(".", 1)[0]) if "." in name else model
            if isinstance(owner, nn.LayerNorm):
                p.fill_(1.0 if leaf == "weight" else 0.0)
            elif isinstance(owner, nn.PReLU):
                p.fill_(0.1 if "feature_fusion" in name else 0.25)
            elif leaf == "relative_position_bias_table":
                t = torch.randn(p.shape, generator=gen)
                bad = t.abs() > 2
                while bad.any():  # rejection sampling == truncated normal on [-2,2]
                    t[bad] = torch.randn(int(bad.sum()), generator=gen)
                    bad = t.abs() > 2
                p.copy_(0.02 * t)
            else:
                if isinstance(owner, nn.ConvTranspose2d):
                    fan_in = owner.in_channels * owner.kernel_size[0] * owner.kernel_size[1]
                elif isinstance(owner, nn.Conv2d):
                    fan_in = (owner.in_channels // owner.groups) * owner.kernel_size[0] * owner.kernel_size[1]
                elif isinstance(owner, nn.Linear):
                    fan_in = owner.in_features
                else:
                    raise RuntimeError(f"unhandled parameter {name}")
                lim = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * lim)
    return model


def build_oracle(seed: int = 0, **cfg) -> OracleBaseModel:
    m = OracleBaseModel(**cfg)
    init_parameters_(m, seed)
    return m.eval()


def init_state_dict(seed: int = 0, **cfg) -> "OrderedDict[str, torch.Tensor]":
    return build_oracle(seed, **cfg).state_dict()


# ----------------------------------------------------------------------------------------------
# homography warp (K1 semantics) -- float64 numpy closed form
# ----------------------------------------------------------------------------------------------
def warp_coords(M: np.ndarray, H: int, W: int):
    """Source coordinates for every destination pixel, float64.  ``M`` maps dst -> src
    (``WARP_INVERSE_MAP``, homography_alignment.py:53); pixel centres at integers."""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3)
    X, Y = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    u = M[0, 0] * X + M[0, 1] * Y + M[0, 2]
    v = M[1, 0] * X + M[1, 1] * Y + M[1, 2]
    w = M[2, 0] * X + M[2, 1] * Y + M[2, 2]
    return u / w, v / w


def warp_frame(src: np.ndarray, M: np.ndarray, quantize_1_32: bool = False) -> np.ndarray:
    """``cv2.warpPerspective(src, M, (W,H), INTER_LINEAR + WARP_INVERSE_MAP)`` semantics
    (homography_alignment.py:46-55,120-129): bilinear taps, taps outside the image contribute 0
    (BORDER_CONSTANT).  ``src`` is ``[H,W,C]``.  A-20: exact coordinates by default; with
    ``quantize_1_32`` source coordinates are rounded to 1/32 px like OpenCV's fixed-point path."""
    src = np.asarray(src)
    H, W, C = src.shape
    sx, sy = warp_coords(M, H, W)
    if quantize_1_32:
        sx = np.rint(sx * 32.0) / 32.0
        sy = np.rint(sy * 32.0) / 32.0
    x0 = np.floor(sx)
    y0 = np.floor(sy)
    ax = sx - x0
    ay = sy - y0
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((H, W, C), dtype=np.float64)
    s64 = src.astype(np.float64)
    for dy, wy in ((0, 1.0 - ay), (1, ay)):
        for dx, wx in ((0, 1.0 - ax), (1, ax)):
            xx, yy = x0 + dx, y0 + dy
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            xs, ys = np.clip(xx, 0, W - 1), np.clip(yy, 0, H - 1)
            out += (wy * wx * ok)[..., None] * s64[ys, xs]
    return out


def warp_burst(burst: np.ndarray, Ms: np.ndarray) -> np.ndarray:
    """``burst [F,H,W,C]``, ``Ms [F,3,3]``; frame 0 is copied, not warped (homography_alignment.py:168,179)."""
    out = np.empty(burst.shape, dtype=np.float64)
    out[0] = burst[0]
    for f in range(1, burst.shape[0]):
        out[f] = warp_frame(burst[f], Ms[f])
    return out


# ----------------------------------------------------------------------------------------------
# full-size tiling -- utils/dataset_utils.py:5-58,140-180 with A-25
# ----------------------------------------------------------------------------------------------
def flow_register_frame(frame: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """``register_frame(frame[H,W,C], flow[H,W,2]) -> [H,W,C]`` of ``registration/optical_flow/register.py:22-47``: every channel is
    sampled at ``mgrid - flow`` (flow's last axis is (dy, dx), ``:32,37``) by ``jsp.ndimage.map_coordinates(order=1,
    mode="nearest")`` (``:11-19``).  Restated from the jax implementation the reference pins (jax 0.4.20,
    ``jax/_src/scipy/ndimage.py``): ``lower = floor(c)``, weights ``(1 - (c - lower), c - lower)``, indices ``lower`` and
    ``lower + 1`` clipped to the image, the four ``weight_y * weight_x * value`` products summed in (lo,lo), (lo,hi), (hi,lo),
    (hi,hi) order; coordinates in float32 as JAX computes them without x64.  Pinned against ``scipy.ndimage.map_coordinates``
    (float64) in ``tests/test_oracle.py``."""
    H, W, _ = frame.shape
    gy, gx = np.mgrid[:H, :W]
    cy = (gy.astype(np.float32) - flow[..., 0].astype(np.float32)).astype(np.float32)
    cx = (gx.astype(np.float32) - flow[..., 1].astype(np.float32)).astype(np.float32)
    ly, lx = np.floor(cy), np.floor(cx)
    uy, ux = (cy - ly).astype(np.float32), (cx - lx).astype(np.float32)
    y0 = np.clip(ly.astype(np.int64), 0, H - 1)
    y1 = np.clip(ly.astype(np.int64) + 1, 0, H - 1)
    x0 = np.clip(lx.astype(np.int64), 0, W - 1)
    x1 = np.clip(lx.astype(np.int64) + 1, 0, W - 1)
    f = frame.astype(np.float32)
    wy0, wx0 = (1 - uy)[..., None], (1 - ux)[..., None]
    wy1, wx1 = uy[..., None], ux[..., None]
    out = ((wy0 * wx0) * f[y0, x0] + (wy0 * wx1) * f[y0, x1]) + (wy1 * wx0) * f[y1, x0]
    return (out + (wy1 * wx1) * f[y1, x1]).astype(np.float32)


def flow_register_burst(burst: np.ndarray, flows: np.ndarray) -> np.ndarray:
    """``[F,H,W,C]`` burst + ``[F-1,H,W,2]`` flows -> registered burst; frame 0 is the reference frame and has no flow
    (``pipeline/real_bsr_iterator.py:121-126``)."""
    out = [burst[0].astype(np.float32)]
    for f in range(1, burst.shape[0]):
        out.append(flow_register_frame(burst[f], flows[f - 1]))
    return np.stack(out)


def tensor_divide_burst(burst: torch.Tensor, psize: int = 80, overlap: int = 40) -> torch.Tensor:
    """``[B,T,C,H,W] -> [B*nh*nw, T, C, psize+2*overlap, psize+2*overlap]`` (tile index row-major,
    batch fastest inside a tile as the reference's ``torch.cat(blocks, 0)`` yields for B=1)."""
    B, T, C, H, W = burst.shape
    hp = (psize - H % psize) % psize
    wp = (psize - W % psize) % psize
    t = burst.reshape(B * T, C, H, W)
    if hp or wp:
        t = F.pad(t, (0, wp, 0, hp), mode="reflect")  # dataset_utils.py:16-33 (A-25 when no pad)
    t = F.pad(t, (overlap, overlap, overlap, overlap), mode="reflect")  # :41-46
    Hp, Wp = H + hp, W + wp
    t = t.view(B, T, C, Hp + 2 * overlap, Wp + 2 * overlap)
    blocks = []
    for i in range(Hp // psize):
        for j in range(Wp // psize):
            blocks.append(t[:, :, :, i * psize : (i + 1) * psize + 2 * overlap, j * psize : (j + 1) * psize + 2 * overlap])
    return torch.cat(blocks, 0)


def tensor_merge(blocks: torch.Tensor, out_hw, psize: int = 320, overlap: int = 160) -> torch.Tensor:
    """dataset_utils.py:140-180 -- centre-crop each tile and stitch; crop to ``out_hw``.  B=1."""
    H, W = out_hw
    Hp = H + (psize - H % psize) % psize
    Wp = W + (psize - W % psize) % psize
    C = blocks.shape[1]
    out = torch.zeros(1, C, Hp, Wp, dtype=blocks.dtype)
    nw = Wp // psize
    for i in range(Hp // psize):
        for j in range(nw):
            part = blocks[i * nw + j]
            out[0, :, i * psize : (i + 1) * psize, j * psize : (j + 1) * psize] = part[:, overlap:-overlap, overlap:-overlap]
    return out[:, :, :H, :W]


def psnr(a: torch.Tensor, b: torch.Tensor, max_val: float = 1.0) -> float:
    """utils/image_utils.py:127 -- ``20 log10(max) - 10 log10(mse)``."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 20.0 * math.log10(max_val) - 10.0 * math.log10(max(mse, 1e-30))


# ------------------------------------------------------------------------------------------------
# 8f-4  ECC homography estimation: `cv2.findTransformECC(template_gray, input_gray, eye(3), MOTION_HOMOGRAPHY, (COUNT|EPS, 100,
# 1e-10))` as called by `homography_alignment.py:19-45` (`register_frame`).  The algorithm lives in OpenCV
# (opencv-contrib-python-headless-rolling 5.0.0.20221015, `pyproject.toml:28`; absent from /root/reference), restated here from
# its published form -- Evangelidis & Psarakis, "Parametric image alignment using enhanced correlation coefficient maximization",
# PAMI 2008, forward-additive scheme, as implemented in `modules/video/src/ecc.cpp` -- and PINNED against the cv2 4.13 in this
# container (`tests/test_oracle.py::test_ecc_oracle_matches_cv2`).  Deliberate difference: the warp-back samples with exact
# bilinear coordinates (OpenCV quantises them to 1/32 px, Appendix A-20), which moves the fixed point by ~1e-3 px.
# ------------------------------------------------------------------------------------------------
def bgr2gray(img: np.ndarray) -> np.ndarray:
    """`cv2.cvtColor(img, COLOR_BGR2GRAY)` on float32 `[H,W,3]` (`homography_alignment.py:39-40`): 0.114 c0 + 0.587 c1 + 0.299 c2."""
    img = img.astype(np.float32)
    return (img[..., 0] * np.float32(0.114) + img[..., 1] * np.float32(0.587) + img[..., 2] * np.float32(0.299)).astype(np.float32)


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def ecc_blur5(img: np.ndarray) -> np.ndarray:
    """`GaussianBlur(img, (5,5), 0)`: OpenCV's fixed small kernel [1,4,6,4,1]/16, separable, BORDER_REFLECT_101."""
    k = np.array([1, 4, 6, 4, 1], np.float64) / 16.0
    H, W = img.shape
    x = img.astype(np.float64)
    cols = _reflect101(np.arange(W)[None, :] + np.arange(-2, 3)[:, None], W)
    x = sum(k[j] * x[:, cols[j]] for j in range(5))
    rows = _reflect101(np.arange(H)[None, :] + np.arange(-2, 3)[:, None], H)
    x = sum(k[j] * x[rows[j], :] for j in range(5))
    return x.astype(np.float32)


def ecc_gradients(img: np.ndarray):
    """`filter2D(img, -1, [-0.5, 0, 0.5])` along x and along y (BORDER_REFLECT_101: the gradient is 0 on the border)."""
    H, W = img.shape
    x = img.astype(np.float64)
    gx = 0.5 * (x[:, _reflect101(np.arange(W) + 1, W)] - x[:, _reflect101(np.arange(W) - 1, W)])
    gy = 0.5 * (x[_reflect101(np.arange(H) + 1, H), :] - x[_reflect101(np.arange(H) - 1, H), :])
    return gx.astype(np.float32), gy.astype(np.float32)


def _bilinear_border0(img: np.ndarray, sx: np.ndarray, sy: np.ndarray) -> np.ndarray:
    H, W = img.shape
    x0, y0 = np.floor(sx), np.floor(sy)
    ax, ay = sx - x0, sy - y0
    x0, y0 = x0.astype(np.int64), y0.astype(np.int64)
    out = np.zeros(sx.shape, np.float64)
    for dy, wy in ((0, 1 - ay), (1, ay)):
        for dx, wx in ((0, 1 - ax), (1, ax)):
            xx, yy = x0 + dx, y0 + dy
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            out += np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0.0) * wy * wx
    return out


def ecc_homography(template: np.ndarray, image: np.ndarray, iters: int = 100, eps: float = 1e-10, warp: np.ndarray = None,
                   quantize_1_32: bool = False):
    """`(rho, warp[3,3]) = findTransformECC(template, image, warp0 = I, MOTION_HOMOGRAPHY)` on single-channel float images of one
    size.  `warp` maps template coordinates to image coordinates (use with WARP_INVERSE_MAP).  float64 arithmetic.
    `quantize_1_32`: round the sampling coordinates to 1/32 px as OpenCV's remap does (test mode: explains the residual
    difference to cv2)."""
    H, W = template.shape
    t = ecc_blur5(template).astype(np.float64)
    im = ecc_blur5(image).astype(np.float64)
    gx, gy = ecc_gradients(im.astype(np.float32))
    gx, gy = gx.astype(np.float64), gy.astype(np.float64)
    M = np.eye(3) if warp is None else np.array(warp, np.float64)
    Y, X = np.mgrid[:H, :W].astype(np.float64)
    rho, last_rho = -1.0, -eps
    for _ in range(iters):
        if abs(rho - last_rho) < eps:
            break
        den = M[2, 0] * X + M[2, 1] * Y + M[2, 2]
        sx = (M[0, 0] * X + M[0, 1] * Y + M[0, 2]) / den
        sy = (M[1, 0] * X + M[1, 1] * Y + M[1, 2]) / den
        if quantize_1_32:
            sx, sy = np.rint(sx * 32.0) / 32.0, np.rint(sy * 32.0) / 32.0
        iw = _bilinear_border0(im, sx, sy)
        gxw = _bilinear_border0(gx, sx, sy)
        gyw = _bilinear_border0(gy, sx, sy)
        rx, ry = np.rint(sx), np.rint(sy)                       # warped all-ones mask, INTER_NEAREST, border 0
        mask = (rx >= 0) & (rx < W) & (ry >= 0) & (ry < H)
        n = mask.sum()
        imean, tmean = iw[mask].mean(), t[mask].mean()
        izm = np.where(mask, iw - imean, iw)                    # subtract(..., mask): pixels outside the mask keep their value
        tzm = np.where(mask, t - tmean, 0.0)
        inorm, tnorm = np.sqrt((izm[mask] ** 2).sum()), np.sqrt((tzm[mask] ** 2).sum())
        # image_jacobian_homo_ECC
        den_ = 1.0 / (X * M[2, 0] + Y * M[2, 1] + 1.0)
        hx = -(X * M[0, 0] + Y * M[0, 1] + M[0, 2]) * den_
        hy = -(X * M[1, 0] + Y * M[1, 1] + M[1, 2]) * den_
        gxp, gyp = gxw * den_, gyw * den_
        tmp = hx * gxp + hy * gyp
        J = np.stack([gxp * X, gyp * X, tmp * X, gxp * Y, gyp * Y, tmp * Y, gxp, gyp], 0).reshape(8, -1)
        Hs = J @ J.T
        Hinv = np.linalg.inv(Hs)
        corr = float((tzm * izm).sum())
        last_rho, rho = rho, corr / (inorm * tnorm)
        ip, tp = J @ izm.reshape(-1), J @ tzm.reshape(-1)
        iph = Hinv @ ip
        lam_n = inorm * inorm - ip @ iph
        lam_d = corr - tp @ iph
        if lam_d <= 0:
            raise ValueError("ECC: the algorithm stopped before its convergence (lambda_d <= 0)")
        lam = lam_n / lam_d
        dp = Hinv @ (lam * tp - ip)                             # J^T (lam * tzm - izm)
        M[0, 0] += dp[0]; M[1, 0] += dp[1]; M[2, 0] += dp[2]
        M[0, 1] += dp[3]; M[1, 1] += dp[4]; M[2, 1] += dp[5]
        M[0, 2] += dp[6]; M[1, 2] += dp[7]
    return rho, M


def homography_coord_diff(Ma, Mb, H: int, W: int) -> float:
    """Largest difference (pixels) between the source coordinates two 3x3 matrices assign to the pixels of an H x W image."""
    Y, X = np.mgrid[:H, :W].astype(np.float64)

    def co(M):
        M = np.asarray(M, np.float64)
        d = M[2, 0] * X + M[2, 1] * Y + M[2, 2]
        return (M[0, 0] * X + M[0, 1] * Y + M[0, 2]) / d, (M[1, 0] * X + M[1, 1] * Y + M[1, 2]) / d
    (a, b), (c, d) = co(Ma), co(Mb)
    return float(max(np.abs(a - c).max(), np.abs(b - d).max()))


# ------------------------------------------------------------------------------------------------
# 8f-3 (first brick): the training loss of train.py.bak:118-119,168 -- CharbonnierLoss + 3 * GWLoss (losses.py:39-80).
# Restated in torch so that autograd supplies the reference gradient; pinned bit-for-bit... to fp32 rounding against vectors produced by
# executing the reference's own losses.py (tests/golden/loss_reference.npz, tests/golden/make_golden_reference.py).
# ------------------------------------------------------------------------------------------------
def charbonnier_loss(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """`CharbonnierLoss(eps)(x, y)` (losses.py:39-51): mean(sqrt((x - y)^2 + eps^2))."""
    d = x - y
    return torch.mean(torch.sqrt(d * d + eps * eps))


def gw_loss(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """`GWLoss()(x1, x2)` (losses.py:53-80): both images clamped to [0,1], per-channel Sobel responses (cross-correlation, zero
    padding), mean((1 + 4|Ix1 - Ix2|)(1 + 4|Iy1 - Iy2|)|x1 - x2|)."""
    x1, x2 = torch.clamp(x1, 0.0, 1.0), torch.clamp(x2, 0.0, 1.0)
    c = x1.shape[1]
    sx = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=x1.dtype).expand(c, 1, 3, 3)
    sy = torch.tensor([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=x1.dtype).expand(c, 1, 3, 3)
    conv = torch.nn.functional.conv2d
    dx = torch.abs(conv(x1, sx, padding=1, groups=c) - conv(x2, sx, padding=1, groups=c))
    dy = torch.abs(conv(x1, sy, padding=1, groups=c) - conv(x2, sy, padding=1, groups=c))
    return torch.mean((1 + 4 * dx) * (1 + 4 * dy) * torch.abs(x1 - x2))


def training_loss(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-3, gw_weight: float = 3.0, clamp_restored: bool = False) -> torch.Tensor:
    """`criterion1(restored, target) + 3 * criterion2(restored, target)` (train.py.bak:168); `clamp_restored`: with the trainer's
    `restored = torch.clamp(restored, 0, 1)` of train.py.bak:167 in front of both criteria."""
    if clamp_restored:
        x = torch.clamp(x, 0.0, 1.0)
    return charbonnier_loss(x, y, eps) + gw_weight * gw_loss(x, y)
