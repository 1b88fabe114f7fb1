"""Arch registry and checkpoint helpers -- the drop-in boundary of the reference
(``/root/reference/fba_net/utils/model_utils.py:1-82``): same function names, argument meaning and
error behaviour, returning the B200-native ``BaseModel``."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

__all__ = ["freeze", "unfreeze", "is_frozen", "save_checkpoint", "load_checkpoint", "load_checkpoint_multigpu",
           "load_start_epoch", "load_optim", "get_arch"]


def freeze(model):
    for p in model.parameters():
        p.requires_grad = False


def unfreeze(model):
    for p in model.parameters():
        p.requires_grad = True


def is_frozen(model):
    return not all(p.requires_grad for p in model.parameters())


def save_checkpoint(model_dir, state, session):
    """``{"epoch", "state_dict", "optimizer"}`` -> ``model_epoch_<epoch>_<session>.pth`` (model_utils.py:22-25)."""
    path = os.path.join(model_dir, "model_epoch_{}_{}.pth".format(state["epoch"], session))
    torch.save(state, path)
    return path


def _strip_module(state_dict):
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def load_checkpoint(model, weights, strict: bool = True):
    """Loads ``checkpoint["state_dict"]``, tolerating the ``module.`` prefix of DataParallel saves
    (model_utils.py:28-38).  Unlike the reference's bare ``except`` it reports what did not match."""
    checkpoint = torch.load(weights, map_location="cpu", weights_only=False)
    target = model.module if isinstance(model, torch.nn.DataParallel) else model
    res = target.load_state_dict(_strip_module(checkpoint["state_dict"]), strict=strict)
    if not strict and (res.missing_keys or res.unexpected_keys):
        print(f"load_checkpoint: missing {list(res.missing_keys)} unexpected {list(res.unexpected_keys)}")
    return res


def load_checkpoint_multigpu(model, weights):
    return load_checkpoint(model, weights)


def load_start_epoch(weights):
    return torch.load(weights, map_location="cpu", weights_only=False)["epoch"]


def load_optim(optimizer, weights):
    checkpoint = torch.load(weights, map_location="cpu", weights_only=False)
    optimizer.load_state_dict(checkpoint["optimizer"])
    lr = None
    for p in optimizer.param_groups:
        lr = p["lr"]
    return lr


def get_arch(opt):
    """``opt.arch == "BaseModel"`` else ``Exception("Arch error!")`` (model_utils.py:65-82).  Reads
    ``train_ps, embed_dim, win_size, token_projection, token_mlp``; optional extras (not in the reference's
    flag set) ``dtype`` ("bf16"/"fp32"), ``in_channels``, ``num_frames``, ``gelu``."""
    from fbanet_b200.model import BaseModel

    arch = opt.arch
    print("You choose " + arch + "...")
    if arch == "BaseModel":
        extra = {k: getattr(opt, k) for k in ("dtype", "in_channels", "num_frames", "gelu") if hasattr(opt, k)}
        model_restoration = BaseModel(
            img_size=opt.train_ps,
            embed_dim=opt.embed_dim,
            window_length=opt.win_size,
            token_projection=opt.token_projection,
            token_mlp=opt.token_mlp,
            **extra,
        )
    else:
        raise Exception("Arch error!")
    return model_restoration
