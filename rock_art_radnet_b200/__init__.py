"""rock_art_radnet_b200 - B200-native (sm_100a CUDA) drop-in for the region-proposal / RoI
hot path of RADNet (Swedish-Rock-Art-Research-Archives/rock-art-radnet).

Reference function            -> here
faster_rcnn/rpn.py   rpn_to_roi, apply_regr_np, non_max_suppression_fast, calc_iou -> .rpn
faster_rcnn/utils.py calc_region_props (= calc_rpn), get_new_img_size, iou          -> .utils
faster_rcnn/RoiPoolingConv.py RoiPoolingConv                                        -> .RoiPoolingConv
batched device pipeline (decode -> sort+NMS -> RoI pool)                            -> .pipeline
per-image sharding + NCCL detection gather                                          -> .sharding

All arithmetic runs in libradnet_b200.so (rock_art_radnet_b200/csrc, C ABI in
include/radnet_b200.h).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .synthetic import HotPathConfig  # noqa: F401

__all__ = ["rpn_to_roi", "apply_regr_np", "non_max_suppression_fast", "calc_iou", "calc_region_props",
           "calc_rpn", "RoiPoolingConv", "ProposalPipeline", "HotPathConfig", "get_new_img_size", "iou"]

_LAZY = {
    "rpn_to_roi": ("rpn", "rpn_to_roi"), "apply_regr_np": ("rpn", "apply_regr_np"),
    "non_max_suppression_fast": ("rpn", "non_max_suppression_fast"), "calc_iou": ("rpn", "calc_iou"),
    "calc_region_props": ("utils", "calc_region_props"), "calc_rpn": ("utils", "calc_rpn"),
    "get_new_img_size": ("utils", "get_new_img_size"), "iou": ("utils", "iou"),
    "RoiPoolingConv": ("RoiPoolingConv", "RoiPoolingConv"),
    "ProposalPipeline": ("pipeline", "ProposalPipeline"),
}


def __getattr__(name):
    # torch is imported only when a compute entry point is first touched
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        obj = getattr(importlib.import_module("." + mod, __name__), attr)
        globals()[name] = obj     # also shadows the same-named submodule (RoiPoolingConv)
        return obj
    raise AttributeError(name)
