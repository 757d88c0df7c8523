"""detectron2 compatibility layer for the hot path.

If a real ``detectron2`` is importable, its registries / structures / config are
used, so ``fsod_train_net.py --eval-only`` and ``demo.py`` pick up the
``CenterNet2Detector`` / ``CenterNet`` / ``CustomCascadeROIHeads`` registered by
``faster_orefsdet_b200.modeling`` unchanged.  Otherwise minimal in-repo stand-ins
of exactly the types the head touches are used (same names, same semantics):

  Registry        fvcore.common.registry.Registry          (d2!/utils/registry.py:5)
  CfgNode         yacs/fvcore CfgNode with _BASE_ + merge_from_list (d2!/config/config.py)
  ShapeSpec       d2!/layers/shape_spec.py
  Boxes           d2!/structures/boxes.py:130-260
  Instances       d2!/structures/instances.py
  ImageList       d2!/structures/image_list.py:70-121
"""
from __future__ import annotations

try:  # pragma: no cover - exercised only where detectron2 is installed
    from detectron2.config import CfgNode  # type: ignore
    from detectron2.layers import ShapeSpec  # type: ignore
    from detectron2.modeling.backbone.build import BACKBONE_REGISTRY  # type: ignore
    from detectron2.modeling.meta_arch.build import META_ARCH_REGISTRY  # type: ignore
    from detectron2.modeling.proposal_generator.build import PROPOSAL_GENERATOR_REGISTRY  # type: ignore
    from detectron2.structures import Boxes, ImageList, Instances  # type: ignore
    from detectron2.utils.registry import Registry  # type: ignore

    HAVE_DETECTRON2 = True
except Exception:  # ImportError or a half-installed detectron2 (missing _C)
    from .config import CfgNode
    from .registry import Registry
    from .structures import Boxes, ImageList, Instances, ShapeSpec

    META_ARCH_REGISTRY = Registry("META_ARCH")
    PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
    BACKBONE_REGISTRY = Registry("BACKBONE")
    HAVE_DETECTRON2 = False

__all__ = ["CfgNode", "Registry", "ShapeSpec", "Boxes", "Instances", "ImageList", "META_ARCH_REGISTRY",
           "PROPOSAL_GENERATOR_REGISTRY", "BACKBONE_REGISTRY", "HAVE_DETECTRON2"]
