"""``FsodRCNN`` registry entry (fewx/modeling/fsod/fsod_rcnn.py:36-37).

The R50-C4 Attention-RPN ancestor shares the skeleton of ``CenterNet2Detector`` and is the
donor of the N-way semantics this package implements (fsod_rcnn.py:472-513), but its own
kernels (1024-channel correlation, polarized channel attention, the global/local/patch
relation of FsodFastRCNNOutputLayers) are a "next" row of the scope table.
The name is registered so configs that select it fail with a clear message instead of a
registry KeyError.
"""
from __future__ import annotations

from torch import nn

from ..compat import register, resolve, META_ARCH_REGISTRY


@register(META_ARCH_REGISTRY)
class FsodRCNN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        raise NotImplementedError(
            "FsodRCNN (R50-C4 Attention-RPN baseline) is not built in this round: the B200 kernels cover the "
            "VoVNet/CenterNet2 path (MODEL.META_ARCHITECTURE=CenterNet2Detector, configs/fsod/finetune_vovnet.yaml).")
