"""Host-side mirror of the reference's modeling interface for the hot path
(fewx/modeling/__init__.py:2, fewx/modeling/fsod/__init__.py:1-6)."""
from ..compat import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, resolve
from .backbone import build_backbone, build_fcos_vovnet_fpn_backbone
from .centernet import CenterNet, CenterNetHead
from .fsod_cen import CenterNet2Detector, PendingBatch
from .fsod_heads import FsodFastRCNNOutputLayers, FsodRes5ROIHeads
from .fsod_rcnn import FsodRCNN
from .resnet import build_resnet_backbone
from .rpn import FsodRPN
from .prototypes import PrototypeBank, SM_Block
from .roi_heads import ROI_HEADS_REGISTRY, CustomCascadeROIHeads, build_roi_heads


def build_model(cfg):
    """d2!/modeling/meta_arch/build.py:16-25."""
    import torch
    model = resolve(META_ARCH_REGISTRY, cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model


__all__ = ["build_model", "build_backbone", "build_roi_heads", "CenterNet", "CenterNetHead", "CenterNet2Detector",
           "FsodRCNN", "FsodRPN", "FsodRes5ROIHeads", "FsodFastRCNNOutputLayers", "build_resnet_backbone", "CustomCascadeROIHeads", "PrototypeBank", "SM_Block", "META_ARCH_REGISTRY",
           "PROPOSAL_GENERATOR_REGISTRY", "BACKBONE_REGISTRY", "ROI_HEADS_REGISTRY"]
