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
    from detectron2.modeling.backbone import Backbone  # type: ignore
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
    from torch.nn import Module as Backbone     # d2!/modeling/backbone/backbone.py: an nn.Module with output_shape()

    META_ARCH_REGISTRY = Registry("META_ARCH")
    PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
    BACKBONE_REGISTRY = Registry("BACKBONE")
    HAVE_DETECTRON2 = False

# ---- registration that can coexist with the reference's own classes ------------------------------------------------
# The reference registers CenterNet2Detector / FsodRCNN / CenterNet under the same names (fewx/modeling/fsod/
# fsod_cen.py:38, fsod_rcnn.py:36, fsod_rpn.py:491) and fvcore registries refuse duplicates.  ``register`` therefore
# never raises: it takes a free name, leaves a taken one alone, and remembers the pair so that
# ``faster_orefsdet_b200.install(override=True)`` can replace the reference's entry explicitly.
_OURS = []          # (registry, name, class)


def register(registry):
    def deco(cls):
        _OURS.append((registry, cls.__name__, cls))
        if cls.__name__ not in registry._obj_map:
            registry._obj_map[cls.__name__] = cls
        return cls
    return deco


def resolve(registry, name: str):
    """The class this package registered under ``name`` (whatever the shared registry currently holds - the reference's
    class of the same name may sit there), else the registry's own entry: the detector always builds ITS sub-modules."""
    for reg, n, cls in _OURS:
        if reg is registry and n == name:
            return cls
    return registry.get(name)


def install_registrations(override: bool = False):
    """Put this package's classes into the registries; with ``override`` also over names that something else (the
    reference's ``fewx``) registered first.  Returns {registry name: {class name: 'installed' | 'kept foreign'}}."""
    report = {}
    for registry, name, cls in _OURS:
        cur = registry._obj_map.get(name)
        if cur is None or cur is cls or override:
            registry._obj_map[name] = cls
            state = "installed"
        else:
            state = "kept foreign"
        report.setdefault(getattr(registry, "_name", "registry"), {})[name] = state
    return report


__all__ = ["CfgNode", "Backbone", "register", "resolve", "install_registrations", "Registry", "ShapeSpec", "Boxes", "Instances", "ImageList", "META_ARCH_REGISTRY",
           "PROPOSAL_GENERATOR_REGISTRY", "BACKBONE_REGISTRY", "HAVE_DETECTRON2"]
