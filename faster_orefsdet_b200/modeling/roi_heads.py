"""Relation ROI head (host side).

Mirrors fewx/modeling/fsod/fsod_roi_heads.py:33-50 (module-local ROI_HEADS_REGISTRY,
``build_roi_heads``) and :282-520 (``CustomCascadeROIHeads``), with the layers the
vendored detectron2 creates for it (d2!/modeling/roi_heads/roi_heads.py:534-592,
cascade_rcnn.py:86-145, box_head.py:66-74, fast_rcnn.py:379-387), under the same
parameter names (log:716-749) so reference checkpoints load:

    conv1, conv2, conv3, fc2, fc3, box_head.0.fc1, box_predictor.0.{cls_score,bbox_pred}

Inference runs three kernels over the whole batch: multi-level ROIAlign, the folded
relation GEMM with its scoring/decoding epilogue, and the class-wise NMS with the
output rescale.  ``fc2`` / ``fc3`` (the 4x4 branch, fsod_roi_heads.py:513-516) never
reach an output in the reference and are parameters only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from .. import fold, ops
from ..compat import register, resolve, Boxes, Instances, Registry, ShapeSpec

ROI_HEADS_REGISTRY = Registry("ROI_HEADS")  # module-local, like fsod_roi_heads.py:33


def build_roi_heads(cfg, input_shape):
    return resolve(ROI_HEADS_REGISTRY, cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


class FastRCNNConvFCHead(nn.Sequential):
    """flatten -> fc1 (FC_DIM / 8 outputs: the fork's d2!/modeling/roi_heads/box_head.py:70) -> ReLU."""

    def __init__(self, cfg, in_channels: int, resolution: int):
        super().__init__()
        h = cfg.MODEL.ROI_BOX_HEAD
        if h.NUM_CONV != 0 or h.NUM_FC != 1:
            raise NotImplementedError("FastRCNNConvFCHead: only NUM_CONV=0, NUM_FC=1 (finetune_vovnet.yaml) is built")
        out = int(h.FC_DIM / 8)
        self.add_module("flatten", nn.Flatten())
        self.add_module("fc1", nn.Linear(in_channels * resolution * resolution, out))
        self.add_module("fc_relu1", nn.ReLU())
        nn.init.kaiming_uniform_(self.fc1.weight, a=1)   # c2_xavier_fill
        nn.init.constant_(self.fc1.bias, 0)
        self.output_size = out


class FastRCNNOutputLayers(nn.Module):
    """cls_score (K+1 logits) and class-agnostic bbox_pred (d2 fast_rcnn.py:379-387)."""

    def __init__(self, cfg, input_size: int):
        super().__init__()
        self.cls_score = nn.Linear(input_size, cfg.MODEL.ROI_HEADS.NUM_CLASSES + 1)
        self.bbox_pred = nn.Linear(input_size, 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        nn.init.constant_(self.cls_score.bias, 0)
        nn.init.constant_(self.bbox_pred.bias, 0)
        self.test_score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
        self.test_nms_thresh = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
        self.test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE


@register(ROI_HEADS_REGISTRY)
class CustomCascadeROIHeads(nn.Module):
    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        r, h, k = cfg.MODEL.ROI_HEADS, cfg.MODEL.ROI_BOX_HEAD, cfg.MODEL.ROI_BOX_CASCADE_HEAD
        self.in_features = self.box_in_features = list(r.IN_FEATURES)
        self.num_classes = r.NUM_CLASSES
        self.strides = [input_shape[f].stride for f in self.in_features]
        channels = {input_shape[f].channels for f in self.in_features}
        if channels != {128} or h.POOLER_RESOLUTION != 8 or h.POOLER_TYPE != "ROIAlignV2" or h.POOLER_SAMPLING_RATIO != 0:
            raise NotImplementedError("CustomCascadeROIHeads: kernels are built for 128-ch maps, 8x8 ROIAlignV2, ratio 0")
        if len(k.IOUS) != 1 or r.NUM_CLASSES != 1 or not h.CLS_AGNOSTIC_BBOX_REG:
            raise NotImplementedError("CustomCascadeROIHeads: one cascade stage, one fg class, class-agnostic boxes")
        self.num_cascade_stages = 1
        self.pooler_resolution, self.pooler_resolution2 = h.POOLER_RESOLUTION, h.POOLER_RESOLUTION2
        self.bbox_reg_weights = tuple(float(x) for x in k.BBOX_REG_WEIGHTS[0])
        self.mult_proposal_score = h.MULT_PROPOSAL_SCORE   # read, but ineffective in the reference (:404 overrides :316)
        self.box_head = nn.ModuleList([FastRCNNConvFCHead(cfg, 128, h.POOLER_RESOLUTION)])
        self.box_predictor = nn.ModuleList([FastRCNNOutputLayers(cfg, self.box_head[0].output_size)])
        # created by the fork's StandardROIHeads.__init__ (head_cnn = True)
        self.fc2 = nn.Linear(2048, 128)
        self.fc3 = nn.Linear(256, 128)
        self.conv1 = nn.Conv2d(128, 64, 1)
        self.conv2 = nn.Conv2d(128, 64, 1)
        self.conv3 = nn.Conv2d(256, 128, 1)
        self._fold_cache = None

    # ------------------------------------------------------------------ folded weights
    def _fold_params(self):
        return [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.conv3.weight, self.conv3.bias,
                self.box_head[0].fc1.weight, self.box_head[0].fc1.bias, self.box_predictor[0].cls_score.weight,
                self.box_predictor[0].cls_score.bias, self.box_predictor[0].bbox_pred.weight,
                self.box_predictor[0].bbox_pred.bias]

    def _state(self) -> Dict[str, torch.Tensor]:
        return {"roi_heads." + k: v.detach() for k, v in self.state_dict().items()}

    def fold_key(self):
        """Identity + version of every parameter the folded matrix and the per-class bias depend on."""
        return tuple((p.data_ptr(), p._version, p.device) for p in self._fold_params())

    def folded(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        key = self.fold_key()
        if self._fold_cache is None or self._fold_cache[0] != key:
            w_fold, w_out, b_out = fold.fold_relation_weights(self._state())
            if w_fold.is_cuda:       # scaled fp16 hi / lo planes for the tensor-core kernel, once per weight load
                w_fold = ops.relation_pack(w_fold)
            self._fold_cache = (key, (w_fold, w_out, b_out))
        return self._fold_cache[1]

    def class_bias(self, support_mean: torch.Tensor) -> torch.Tensor:
        """support_mean [C,128,8,8] -> per-class folded bias [C,128]."""
        return fold.fold_class_bias(self._state(), support_mean)

    # ------------------------------------------------------------------ raw batched path
    @torch.no_grad()
    def detect_raw(self, features: Sequence[torch.Tensor], bias_cls: torch.Tensor, rois: torch.Tensor,
                   roi_count: torch.Tensor, num_classes: int, image_hw: torch.Tensor, out_hw: Optional[torch.Tensor],
                   status: torch.Tensor, feature_bounds: Optional[Sequence[torch.Tensor]] = None):
        """features[l] [B,128,H,W] raw backbone maps; rois [B*C,cap,4]; returns the padded
        outputs of ops.final_detect plus the per-ROI (boxes, scores).  ``feature_bounds[l]``: device scalar bounding
        max|features[l]|, or [B] floats with one bound per image, when the producer reported it (the FPN output
        convolutions do); computed otherwise."""
        w_fold, w_out, b_out = self.folded()
        pooled = ops.roi_align(features, self.strides, rois, roi_count, num_classes, self.pooler_resolution, tiled=True)
        B = features[0].shape[0]
        if feature_bounds is None or any(b is None for b in feature_bounds):
            feature_bounds = [f.abs().amax((1, 2, 3)) for f in features]          # generic path: per-image max|map|
        # [levels, B]: the rows of image b are scaled by ITS maps' bounds (a [1]-shaped bound is shared by the batch)
        x_amax = torch.stack([b.reshape(-1).expand(B) for b in feature_bounds]).contiguous()
        det_boxes, det_scores = ops.relation_head(pooled, w_fold, bias_cls, w_out, b_out, rois, roi_count, num_classes,
                                                  self.bbox_reg_weights, x_amax=x_amax)
        p = self.box_predictor[0]
        out = ops.final_detect(det_boxes, det_scores, roi_count, num_classes, p.test_score_thresh, p.test_nms_thresh,
                               p.test_topk_per_image, image_hw, out_hw, status)
        return out, (det_boxes, det_scores)

    # ------------------------------------------------------------------ reference-shaped forward
    def forward(self, images, features: Dict[str, torch.Tensor], support_box_features: List[torch.Tensor],
                proposals: List[Instances], targets=None):
        """(fsod_roi_heads.py:374-401) one support class: support_box_features = [rcnn_8 [S,128,8,8], rcnn_4 [S,128,4,4]]."""
        if self.training:
            raise NotImplementedError("CustomCascadeROIHeads: training is outside the inference hot path")
        del images
        feats = [features[f] for f in self.box_in_features]
        dev = feats[0].device
        B = len(proposals)
        cap = max(max((len(p) for p in proposals), default=1), 1)
        rois = torch.zeros((B, cap, 4), dtype=torch.float32, device=dev)
        counts = []
        for i, p in enumerate(proposals):
            n = len(p)
            counts.append(n)
            if n:
                rois[i, :n] = p.proposal_boxes.tensor
        roi_count = torch.tensor(counts, dtype=torch.int32, device=dev)
        image_hw = torch.tensor([list(p.image_size) for p in proposals], dtype=torch.int32, device=dev)
        bias = self.class_bias(support_box_features[0].to(dev).mean(0, True))
        status = ops.new_status(dev)
        (ob, os_, ocls, _, oc), _ = self.detect_raw(feats, bias, rois, roi_count, 1, image_hw, None, status)
        ops.check_status(status)
        return pack_instances(ob, os_, ocls, oc, [p.image_size for p in proposals]), {}


LAST_D2H_BYTES = 0     # size of the last detections transfer (bench.py reports it)


def pack_block(boxes, scores, classes, count) -> torch.Tensor:
    """Padded device outputs -> one [B, K, 7] fp32 block (box, score, class, count) that leaves in ONE transfer."""
    B, K = scores.shape
    return torch.cat((boxes, scores.unsqueeze(-1), classes.to(torch.float32).unsqueeze(-1),
                      count.to(torch.float32).view(B, 1, 1).expand(B, K, 1)), -1)


def _targets_cpu(args, kwargs) -> bool:
    dev = kwargs.get("device", args[0] if args else None)
    if set(kwargs) - {"device", "non_blocking"} or len(args) > 1 or dev is None:
        return False
    return (isinstance(dev, str) and dev == "cpu") or (isinstance(dev, torch.device) and dev.type == "cpu")


def _field_stamp(v):
    t = v.tensor if isinstance(v, Boxes) else v
    return (v, t._version if isinstance(t, torch.Tensor) else None)


class DetectionInstances(Instances):
    """``Instances`` (detectron2's own class when it is installed, the compat stand-in otherwise) whose fields also exist
    on the host: the detector moves all detections of a batch with ONE device-to-host transfer and attaches the host
    views.  ``.to("cpu")`` - what COCOEvaluator.process does per image (fewx/evaluation/coco_evaluation.py:119-126) -
    returns them without touching the device, as long as no field was added, removed, replaced or modified in place
    since construction; otherwise (and for every other target) it is the base class's field-by-field ``to``."""

    def to(self, *args, **kwargs):
        mirror = self.__dict__.get("_host_mirror")
        if mirror is not None and _targets_cpu(args, kwargs):
            stamps, fields = self.__dict__["_mirror_stamps"], self._fields
            if len(fields) == len(stamps) and all(k in fields and fields[k] is st[0] and _field_stamp(fields[k])[1] == st[1]
                                                  for k, st in stamps.items()):
                ret = Instances(self._image_size)
                for k, v in mirror.items():
                    ret.set(k, v)
                return ret
        return super().to(*args, **kwargs)


def mirrored_instances(image_size, fields, host_fields) -> "DetectionInstances":
    """DetectionInstances over ``fields`` (device) with ``host_fields`` as their completed host copies; the fields must
    have equal lengths (not re-checked: the hot path builds one of these per image)."""
    inst = DetectionInstances.__new__(DetectionInstances)
    inst.__dict__["_image_size"] = image_size
    inst.__dict__["_fields"] = fields
    inst.__dict__["_host_mirror"] = host_fields
    inst.__dict__["_mirror_stamps"] = {k: _field_stamp(v) for k, v in fields.items()}
    return inst


def pack_instances(boxes, scores, classes, count, image_sizes) -> List[Instances]:
    """Padded device outputs -> list[Instances] on the device, with ONE device-to-host transfer of the whole
    padded block ([B, K, 7] fp32) whose views ride along as the host mirror of every DetectionInstances
    (its .to("cpu") returns them; the reference's evaluator copies field by field, image by image:
    fewx/evaluation/coco_evaluation.py:119-126).  The valid rows are compacted once and split per image with
    split_with_sizes, so the per-image Python cost is the construction of the containers only."""
    block = pack_block(boxes, scores, classes, count)
    host = torch.empty(block.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(block, non_blocking=True)
    torch.cuda.current_stream(scores.device).synchronize()
    return instances_from_block(block, host, image_sizes)


def instances_from_block(block: torch.Tensor, host: torch.Tensor, image_sizes) -> List[Instances]:
    """``block``: the device block of pack_block, ``host``: its (completed) pinned host copy."""
    B, K = host.shape[0], host.shape[1]
    global LAST_D2H_BYTES
    LAST_D2H_BYTES = host.numel() * 4
    counts_t = host[:, 0, 6].to(torch.int64)
    counts = counts_t.tolist()
    idx = torch.nonzero((torch.arange(K)[None, :] < counts_t[:, None]).flatten()).squeeze(1)
    flat_h = host.view(B * K, 7).index_select(0, idx)
    # (pinned: the copy of the index list must not wait for whatever the stream is running - a later batch, say)
    flat_d = block.view(B * K, 7).index_select(0, idx.pin_memory().to(block.device, non_blocking=True))
    parts = []
    for flat in (flat_d, flat_h):
        parts.append((torch.split(flat[:, :4], counts), torch.split(flat[:, 4], counts),
                      torch.split(flat[:, 5].to(torch.int64), counts)))
    (db, ds, dc), (hb, hs, hc) = parts
    def boxes_of(t):          # the fields are already [n, 4] fp32: skip the constructor's conversions and checks
        bx = Boxes.__new__(Boxes)
        bx.tensor = t
        return bx

    out = []
    for b in range(B):
        out.append(mirrored_instances((int(image_sizes[b][0]), int(image_sizes[b][1])),
                                      {"pred_boxes": boxes_of(db[b]), "scores": ds[b], "pred_classes": dc[b]},
                                      {"pred_boxes": boxes_of(hb[b]), "scores": hs[b], "pred_classes": hc[b]}))
    return out
