"""B200-native support-guided detection head of Faster-OreFSDet.

    from faster_orefsdet_b200.config import get_cfg
    from faster_orefsdet_b200.modeling import build_model

The hot path runs in hand-written sm_100a kernels behind a C ABI
(include/fod_b200.h -> faster_orefsdet_b200/libfod_b200.so); see DESIGN.md.
"""
__version__ = "0.1.0"


def install(override: bool = False):
    """Register this package's detector under the reference's registry names.

    ``import faster_orefsdet_b200.modeling`` alone takes every FREE name and never raises.  When the reference's
    ``fewx`` has been imported as well (``fsod_train_net.py:18`` does ``from fewx.config import get_cfg``, which pulls in
    ``fewx/__init__.py:1`` -> ``fewx.modeling`` -> its own ``@META_ARCH_REGISTRY.register()`` classes of the same names),
    call ``install(override=True)`` AFTER that import: it replaces ``META_ARCH_REGISTRY["CenterNet2Detector"]``,
    ``["FsodRCNN"]``, ``PROPOSAL_GENERATOR_REGISTRY["CenterNet"]``, ``BACKBONE_REGISTRY["build_fcos_vovnet_fpn_backbone"]``
    and the reference's module-local ``fewx.modeling.fsod.fsod_roi_heads.ROI_HEADS_REGISTRY["CustomCascadeROIHeads"]``
    (fsod_roi_heads.py:33,45-50), so ``detectron2.modeling.build_model(cfg)`` (d2!/modeling/meta_arch/build.py:16-25) -
    what ``fsod_train_net.py --eval-only`` and ``demo.py`` call - builds the B200 detector.  Returns a report dict."""
    import sys

    from . import modeling  # noqa: F401  (defines and registers the classes)
    from .compat import install_registrations
    report = install_registrations(override)
    ref = sys.modules.get("fewx.modeling.fsod.fsod_roi_heads")
    if ref is not None and hasattr(ref, "ROI_HEADS_REGISTRY"):
        from .modeling.roi_heads import CustomCascadeROIHeads
        cur = ref.ROI_HEADS_REGISTRY._obj_map.get("CustomCascadeROIHeads")
        if cur is None or override:
            ref.ROI_HEADS_REGISTRY._obj_map["CustomCascadeROIHeads"] = CustomCascadeROIHeads
            report.setdefault("fewx ROI_HEADS", {})["CustomCascadeROIHeads"] = "installed"
        else:
            report.setdefault("fewx ROI_HEADS", {})["CustomCascadeROIHeads"] = "kept foreign"
    return report
