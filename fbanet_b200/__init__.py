"""fbanet_b200 -- B200-native (sm_100a) implementation of the FBANet ``BaseModel`` burst-SR forward path.

Public API (mirrors the reference's): ``get_arch(opt)``, ``BaseModel``, ``load_checkpoint``; the alignment front end
(``ecc_homography_burst`` = ``cv2.findTransformECC``, ``warp_burst`` = ``cv2.warpPerspective``, ``flow_warp_burst`` = the optical-flow
``register_frame``); the full-size tiled drivers in ``fbanet_b200.tiling`` (single GPU and row-band sharded over the GPUs of one box);
and the training step in ``fbanet_b200.train`` (``train_step``: training-mode forward on a reverse-mode tape over the C-ABI ops, ``training_loss``,
backward into ``FlatParams`` flat buffers, bucketed gradient all-reduce, fused AdamW; learning-rate schedules and stochastic-depth rates).
The per-kernel ops are also registered as ``torch.library`` custom ops (``torch.ops.fbanet.*``, :mod:`fbanet_b200.torch_ops`).
"""
from . import torch_ops  # noqa: F401  (registers torch.ops.fbanet.*)
from .model import BaseModel, HostResult  # noqa: F401
from .ops import ecc_homography_burst, flow_warp_burst, training_loss, warp_burst  # noqa: F401
from .utils.model_utils import get_arch, load_checkpoint, load_checkpoint_multigpu, load_optim, load_start_epoch, save_checkpoint  # noqa: F401

__all__ = ["BaseModel", "HostResult", "get_arch", "load_checkpoint", "load_checkpoint_multigpu", "load_start_epoch", "load_optim", "save_checkpoint", "warp_burst",
           "flow_warp_burst", "ecc_homography_burst", "training_loss"]
