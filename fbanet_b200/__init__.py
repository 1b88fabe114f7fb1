"""fbanet_b200 -- B200-native (sm_100a) implementation of the FBANet ``BaseModel`` burst-SR forward path.

Public API (mirrors the reference's): ``get_arch(opt)``, ``BaseModel``, ``load_checkpoint``; plus
``warp_burst`` (homography warp front end) and the full-size tiled driver in ``fbanet_b200.tiling``.
"""
from .model import BaseModel  # noqa: F401
from .ops import warp_burst  # noqa: F401
from .utils.model_utils import get_arch, load_checkpoint, load_checkpoint_multigpu, load_optim, load_start_epoch, save_checkpoint  # noqa: F401

__all__ = ["BaseModel", "get_arch", "load_checkpoint", "load_checkpoint_multigpu", "load_start_epoch", "load_optim", "save_checkpoint", "warp_burst"]
