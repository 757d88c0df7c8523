"""CPU oracle for the support-guided detection head of Faster-OreFSDet.

TEST INFRASTRUCTURE ONLY.  Nothing in ``faster_orefsdet_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the CPU
arm, never as the product path.

What it is: a function-for-function fp32 restatement, in plain PyTorch-CPU +
torchvision-CPU ops, of the reference's ``MODEL.DEVICE=cpu`` inference path for
``configs/fsod/finetune_vovnet.yaml`` (SURVEY.md section 8a, rows Q1..O1, P1/P2,
N1).  Each function cites the reference file:line it follows.  Paths are
relative to /root/reference; ``d2!/`` means inside ``detectron2.7z`` (the
vendored, modified detectron2 v0.5).

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4).  The oracle is pinned instead against the *reference's own
code executed in the build container*: ``tests/golden/make_golden.py`` imports the
real ``fewx`` / vendored ``detectron2`` modules (with small stand-ins for the
third-party packages that are not installed) and records input/output vectors
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through
this file.  See DESIGN.md "Oracle".

Third-party arithmetic the reference delegates to (not under /root/reference):
torchvision (pin 0.8.2+cu101, here 0.26.0) ``ops.nms`` / ``ops.roi_align`` /
``ops.boxes.batched_nms``; ATen ``conv2d``, ``group_norm``, ``topk``,
``kthvalue``.  The same entry points (CPU) are called here.

Tie rules (the reference leaves these implementation-defined; the oracle fixes
them, the CUDA path follows them, the tests engineer them):
  * pre-NMS top-k (fsod_rpn.py:1157-1162, ``topk(sorted=False)``): the selected
    set is every score above the k-th largest plus the lowest-location-index
    ties at it; survivors are emitted in ascending location index.
  * NMS order: stable descending sort (lower input index first among equal
    scores) - torchvision's CPU kernel since 0.9; 0.8.2 was unstable.
  * post-NMS top-k (fsod_rpn.py:1198-1206): ``score >= kth`` keeps every tie.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
import torchvision
from torchvision.ops import roi_align as tv_roi_align
from torchvision.ops import nms as tv_nms

Tensor = torch.Tensor

LEVELS = ("p3", "p4", "p5")
STRIDES = (8, 16, 32)
SCALE_CLAMP = math.log(1000.0 / 16)  # d2!/modeling/box_regression.py:13


# --------------------------------------------------------------------------- #
# configuration actually read by the path (resolved log:39-545)
# --------------------------------------------------------------------------- #
class HeadConfig:
    """Resolved values of the config keys the hot path reads
    (log/fsod_finetune_stone_vovnet_25_test_log.txt:199-249, 343-397, 534)."""

    def __init__(self, **kw):
        self.inference_th = 1e-5          # MODEL.CENTERNET.INFERENCE_TH
        self.pre_nms_topk = 1000          # MODEL.CENTERNET.PRE_NMS_TOPK_TEST
        self.post_nms_topk = 256          # MODEL.CENTERNET.POST_NMS_TOPK_TEST
        self.nms_th = 0.6                 # MODEL.CENTERNET.NMS_TH_TEST
        self.strides = STRIDES            # MODEL.CENTERNET.FPN_STRIDES
        self.pooler_resolution = 8        # MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        self.pooler_resolution2 = 4       # MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION2
        self.bbox_reg_weights = (10.0, 10.0, 5.0, 5.0)  # ROI_BOX_CASCADE_HEAD.BBOX_REG_WEIGHTS[0]
        self.score_thresh_test = 0.0      # MODEL.ROI_HEADS.SCORE_THRESH_TEST (fork default)
        self.nms_thresh_test = 0.9        # MODEL.ROI_HEADS.NMS_THRESH_TEST
        self.detections_per_image = 100   # TEST.DETECTIONS_PER_IMAGE
        self.gn_groups = 32
        self.gn_eps = 1e-5
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


# --------------------------------------------------------------------------- #
# Q1: support taps                                  fsod_cen.py:458-460,476-479,498-500
# --------------------------------------------------------------------------- #
def support_taps(proto: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """proto [1,128,h,w] -> k11 [128], k13 [128,3], k31 [128,3]
    (AdaptiveAvgPool2d (1,1), (1,3), (3,1); fsod_cen.py:72-75)."""
    k11 = F.adaptive_avg_pool2d(proto, (1, 1))[0, :, 0, 0]
    k13 = F.adaptive_avg_pool2d(proto, (1, 3))[0, :, 0, :]
    k31 = F.adaptive_avg_pool2d(proto, (3, 1))[0, :, :, 0]
    return k11.contiguous(), k13.contiguous(), k31.contiguous()


# --------------------------------------------------------------------------- #
# Q2 + Q3: depthwise correlation and the 1x1 relation conv on maps
#                                                    fsod_cen.py:463-470 (p3), 482-491, 502-509
# --------------------------------------------------------------------------- #
def correlate_level(q: Tensor, k11: Tensor, k13: Tensor, k31: Tensor,
                    w3: Tensor, b3: Tensor) -> Tensor:
    """q [B,128,H,W]; taps as from support_taps; w3 [128,256,1,1], b3 [128].
    Returns attn [B,128,H,W]."""
    C = q.shape[1]
    a = F.relu(F.conv2d(q, k11.view(C, 1, 1, 1), groups=C))
    a = F.relu(F.conv2d(a, k11.view(C, 1, 1, 1), groups=C))
    b = F.relu(F.conv2d(q, k13.view(C, 1, 1, 3), padding=(0, 1), groups=C))
    b = F.relu(F.conv2d(b, k31.view(C, 1, 3, 1), padding=(1, 0), groups=C))
    s = a + b + q
    return F.relu(F.conv2d(torch.cat((s, q), 1), w3, b3))


# --------------------------------------------------------------------------- #
# H0: CenterNetHead (boundary; stays cuDNN in the product)
#                    CenterNet2/centernet/modeling/dense_heads/centernet_head.py:141-161
# --------------------------------------------------------------------------- #
def centernet_head_level(attn: Tensor, sd: Dict[str, Tensor], level: int,
                         cfg: HeadConfig, prefix="proposal_generator.centernet_head.") -> Tuple[Tensor, Tensor]:
    """Returns (agn_hm logits [B,1,H,W], reg = relu(scale_l * bbox_pred) [B,4,H,W])."""
    p = prefix
    t = F.conv2d(attn, sd[p + "bbox_tower.0.weight"], sd[p + "bbox_tower.0.bias"], padding=1)
    t = F.group_norm(t, cfg.gn_groups, sd[p + "bbox_tower.1.weight"], sd[p + "bbox_tower.1.bias"], cfg.gn_eps)
    t = F.relu(t)
    hm = F.conv2d(t, sd[p + "agn_hm.weight"], sd[p + "agn_hm.bias"], padding=1)
    reg = F.conv2d(t, sd[p + "bbox_pred.weight"], sd[p + "bbox_pred.bias"], padding=1)
    reg = F.relu(reg * sd[p + f"scales.{level}.scale"])
    return hm, reg


# --------------------------------------------------------------------------- #
# D1/D2: sigmoid, threshold, top-k, box decode      fsod_rpn.py:782-800, 1071-1074, 1116-1181
# --------------------------------------------------------------------------- #
def decode_level(hm_logits: Tensor, reg: Tensor, stride: int, cfg: HeadConfig, is_logit: bool = True
                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """One image, one level.  hm_logits [H,W] (pre-sigmoid), reg [4,H,W]
    (post relu*scale, NOT yet multiplied by stride; fsod_rpn.py:1107 does that).
    Returns (loc int64 [n] ascending, boxes [n,4], scores [n])."""
    H, W = hm_logits.shape
    p = (torch.sigmoid(hm_logits) if is_logit else hm_logits).reshape(-1)   # :1073, :1126-1127
    cand = torch.nonzero(p > cfg.inference_th).squeeze(1)        # :1131, :1147
    k = min(int(cand.numel()), cfg.pre_nms_topk)                 # :1132-1134
    if cand.numel() > k:                                         # :1157
        # top-k with the oracle's tie rule (see module docstring)
        order = torch.sort(p[cand], descending=True, stable=True).indices[:k]
        cand = torch.sort(cand[order]).values
    ys = torch.div(cand, W, rounding_mode="floor")
    xs = cand - ys * W
    gx = (xs * stride).to(torch.float32) + (stride // 2)         # :786-798
    gy = (ys * stride).to(torch.float32) + (stride // 2)
    r = (reg * float(stride)).reshape(4, -1)[:, cand]            # :1107, :1128-1129
    x1 = gx - r[0]
    y1 = gy - r[1]
    x2 = gx + r[2]
    y2 = gy + r[3]
    x2 = torch.max(x2, x1 + 0.01)                                # :1172-1173
    y2 = torch.max(y2, y1 + 0.01)
    boxes = torch.stack((x1, y1, x2, y2), 1)
    scores = torch.sqrt(p[cand])                                 # :1175 (WITH_AGN_HM)
    return cand, boxes, scores


# --------------------------------------------------------------------------- #
# N0: NMS + post-NMS top-k                           fsod_rpn.py:1184-1210; ml_nms.py:4-31;
#                                                    d2!/layers/nms.py:10-30
# --------------------------------------------------------------------------- #
def batched_nms_coordinate_trick(boxes: Tensor, scores: Tensor, idxs: Tensor, thr: float) -> Tensor:
    """torchvision 0.8.2 ``boxes.batched_nms`` (the reference's pin, log:19): offset
    every box by ``idx * (max_coordinate + 1)`` in fp32, then one plain NMS.
    torchvision.ops.nms (CPU): stable descending sort, suppress when
    ``inter / (area_i + area_j - inter) > thr`` (fp32 ratio compared in double)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    return tv_nms(boxes + offsets[:, None], scores, thr)


def proposal_nms_topk(boxes: Tensor, scores: Tensor, cfg: HeadConfig) -> Tensor:
    """Returns indices (into the concatenated per-image candidate list) of the
    proposals handed to the ROI head, score-descending."""
    keep = batched_nms_coordinate_trick(boxes, scores, torch.zeros_like(scores, dtype=torch.int64), cfg.nms_th)
    n = keep.numel()
    if n > cfg.post_nms_topk:                                    # :1198
        s = scores[keep]
        thr = torch.kthvalue(s, n - cfg.post_nms_topk + 1).values   # :1200-1203
        keep = keep[torch.nonzero(s >= thr).squeeze(1)]          # :1204-1206
    return keep


# --------------------------------------------------------------------------- #
# R1: level assignment + ROIAlign                    d2!/modeling/poolers.py:22-58,190-250;
#                                                    d2!/layers/roi_align.py:49-65
# --------------------------------------------------------------------------- #
def assign_levels(boxes: Tensor, min_level=3, max_level=5, canonical_size=224, canonical_level=4) -> Tensor:
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])     # d2!/structures/boxes.py:181-190
    lvl = torch.floor(canonical_level + torch.log2(torch.sqrt(area) / canonical_size + 1e-8))
    lvl = torch.clamp(lvl, min=min_level, max=max_level)
    return lvl.to(torch.int64) - min_level


def roi_pool(feats: Sequence[Tensor], boxes_per_image: Sequence[Tensor], out_size: int,
             strides: Sequence[int] = STRIDES) -> Tensor:
    """ROIPooler.forward with ROIAlignV2 (aligned=True), sampling_ratio=0."""
    rois = torch.cat([torch.cat((torch.full_like(b[:, :1], float(i)), b), 1)
                      for i, b in enumerate(boxes_per_image)], 0)
    lv = assign_levels(torch.cat(list(boxes_per_image), 0))
    out = torch.zeros((rois.shape[0], feats[0].shape[1], out_size, out_size), dtype=feats[0].dtype)
    for l, f in enumerate(feats):
        idx = torch.nonzero(lv == l).squeeze(1)
        out[idx] = tv_roi_align(f, rois[idx], (out_size, out_size), 1.0 / strides[l], 0, True)
    return out


# --------------------------------------------------------------------------- #
# R2: relation head                                   fsod_roi_heads.py:482-483,509-511,520
# --------------------------------------------------------------------------- #
def relation_head(x: Tensor, sup8: Tensor, sd: Dict[str, Tensor], prefix="roi_heads.") -> Tuple[Tensor, Tensor]:
    """x [R,128,8,8] pooled proposals, sup8 [S,128,8,8] pooled support boxes.
    Returns (logits [R,2], deltas [R,4]).  The 4x4 branch (fsod_roi_heads.py:513-516)
    never reaches an output and is omitted."""
    p = prefix
    s = sup8.mean(0, True).expand_as(x)
    a = F.conv2d(torch.cat((x, s), 1), sd[p + "conv3.weight"], sd[p + "conv3.bias"]) + torch.cat(
        (F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"]),
         F.conv2d(s, sd[p + "conv2.weight"], sd[p + "conv2.bias"])), 1)
    f = F.relu(F.linear(a.flatten(1), sd[p + "box_head.0.fc1.weight"], sd[p + "box_head.0.fc1.bias"]))
    logits = F.linear(f, sd[p + "box_predictor.0.cls_score.weight"], sd[p + "box_predictor.0.cls_score.bias"])
    deltas = F.linear(f, sd[p + "box_predictor.0.bbox_pred.weight"], sd[p + "box_predictor.0.bbox_pred.bias"])
    return logits, deltas


# --------------------------------------------------------------------------- #
# R3: softmax + apply_deltas                          custom_fast_rcnn.py:160-170;
#                                                    d2!/modeling/box_regression.py:77-115
# --------------------------------------------------------------------------- #
def apply_deltas(deltas: Tensor, boxes: Tensor, weights=(10.0, 10.0, 5.0, 5.0)) -> Tensor:
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    dx = deltas[:, 0] / weights[0]
    dy = deltas[:, 1] / weights[1]
    dw = torch.clamp(deltas[:, 2] / weights[2], max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3] / weights[3], max=SCALE_CLAMP)
    pcx = dx * w + cx
    pcy = dy * h + cy
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), 1)


def score_and_decode(logits: Tensor, deltas: Tensor, proposals: Tensor, cfg: HeadConfig) -> Tuple[Tensor, Tensor]:
    """Returns (fg score [R], boxes [R,4]).  One cascade stage (log:343-350), so the
    stage average is the identity; MULT_PROPOSAL_SCORE has no effect because the
    second ``_forward_box`` definition wins (fsod_roi_heads.py:404 over :316)."""
    probs = F.softmax(logits, dim=-1)
    return probs[:, 0], apply_deltas(deltas, proposals, cfg.bbox_reg_weights)


# --------------------------------------------------------------------------- #
# R4 / N1: final class-wise NMS                       d2!/modeling/roi_heads/fast_rcnn.py:118-171;
#                                                    fsod_fast_rcnn.py:84-145
# --------------------------------------------------------------------------- #
def clip_boxes(boxes: Tensor, image_size: Tuple[int, int]) -> Tensor:
    h, w = image_size                                            # d2!/structures/boxes.py:192-207
    return torch.stack((boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h),
                        boxes[:, 2].clamp(0, w), boxes[:, 3].clamp(0, h)), 1)


def final_detect(boxes: Tensor, scores: Tensor, class_idx: Tensor, image_size: Tuple[int, int],
                 cfg: HeadConfig) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """boxes [N,4], scores [N], class_idx [N] int64 (position of the class in the
    episode's class list).  Returns (boxes, scores, class_idx, kept row index)."""
    valid = torch.isfinite(boxes).all(1) & torch.isfinite(scores)
    rows = torch.nonzero(valid).squeeze(1)
    boxes, scores, class_idx = clip_boxes(boxes[rows], image_size), scores[rows], class_idx[rows]
    m = torch.nonzero(scores > cfg.score_thresh_test).squeeze(1)
    boxes, scores, class_idx, rows = boxes[m], scores[m], class_idx[m], rows[m]
    keep = batched_nms_coordinate_trick(boxes, scores, class_idx, cfg.nms_thresh_test)
    keep = keep[: cfg.detections_per_image]
    return boxes[keep], scores[keep], class_idx[keep], rows[keep]


# --------------------------------------------------------------------------- #
# O1: detector_postprocess                            d2!/modeling/postprocessing.py:9-75
# --------------------------------------------------------------------------- #
def postprocess(boxes: Tensor, scores: Tensor, classes: Tensor, image_size: Tuple[int, int],
                out_h: int, out_w: int) -> Tuple[Tensor, Tensor, Tensor]:
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    b = boxes.clone()
    b[:, 0::2] *= sx
    b[:, 1::2] *= sy
    b = clip_boxes(b, (out_h, out_w))
    ne = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)      # d2!/structures/boxes.py:209-222
    return b[ne], scores[ne], classes[ne]


# --------------------------------------------------------------------------- #
# P1 / P2: prototype builder                           fsod_cen.py:348-389, 584-630
# --------------------------------------------------------------------------- #
def sm_block(x: Tensor, sd: Dict[str, Tensor], prefix: str, seg_dim: int) -> Tensor:
    """SM_Block.forward (eval: dropouts are identity).  x [B,H,W,C]."""
    B, H, W, C = x.shape
    S = C // seg_dim
    h = x.reshape(B, H, W, seg_dim, S).permute(0, 3, 2, 1, 4).reshape(B, seg_dim, W, H * S)
    h = F.linear(h, sd[prefix + "mlp_h.weight"])
    h = h.reshape(B, seg_dim, W, H, S).permute(0, 3, 2, 1, 4).reshape(B, H, W, C)
    w = x.reshape(B, H, W, seg_dim, S).permute(0, 3, 1, 2, 4).reshape(B, seg_dim, H, W * S)
    w = F.linear(w, sd[prefix + "mlp_w.weight"])
    w = w.reshape(B, seg_dim, H, W, S).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
    a = (h + w).permute(0, 3, 1, 2).flatten(2).mean(2)
    a = F.linear(a, sd[prefix + "reweighting.fc1.weight"], sd[prefix + "reweighting.fc1.bias"])
    a = F.linear(F.gelu(a), sd[prefix + "reweighting.fc2.weight"], sd[prefix + "reweighting.fc2.bias"])
    a = a.reshape(B, C, 2).permute(2, 0, 1).softmax(0).unsqueeze(2).unsqueeze(2)
    y = w * a[0] + h * a[1]
    return F.linear(y, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def build_prototypes(support_feats: Dict[str, Tensor], support_boxes: Sequence[Tensor],
                     sd: Dict[str, Tensor], cfg: HeadConfig) -> Dict[str, Tensor]:
    """One class.  support_feats[l] [S,128,h_l,w_l] backbone maps of the S support
    images, support_boxes: S tensors [1,4].  Returns the five pkl entries of
    ``support_feature.pkl`` for that class (fsod_cen.py:384-389)."""
    feats = [support_feats[l] for l in LEVELS]
    out = {
        "rcnn_8": roi_pool(feats, support_boxes, cfg.pooler_resolution, cfg.strides),
        "rcnn_4": roi_pool(feats, support_boxes, cfg.pooler_resolution2, cfg.strides),
    }
    for l, size in zip(LEVELS, (32, 16, 8)):
        x = F.adaptive_avg_pool2d(support_feats[l], (size, size)).permute(0, 2, 3, 1)
        y = sm_block(x, sd, f"vip_{l}.", size).permute(0, 3, 2, 1)
        out[l] = y.mean(0, True)
    return out


# --------------------------------------------------------------------------- #
# whole head for one query image (reference asserts B == 1, fsod_cen.py:438)
# --------------------------------------------------------------------------- #
def propose(features: Dict[str, Tensor], proto: Dict[str, Tensor], sd: Dict[str, Tensor],
            cfg: HeadConfig, trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """Rows Q1..N0 for ONE image (features[l] is [1,128,H,W]) and ONE class.
    Returns (proposal_boxes [R,4], objectness [R])."""
    boxes, scores = [], []
    for li, l in enumerate(LEVELS):
        k11, k13, k31 = support_taps(proto[l])
        attn = correlate_level(features[l], k11, k13, k31, sd["conv3.weight"], sd["conv3.bias"])
        hm, reg = centernet_head_level(attn, sd, li, cfg)
        loc, b, s = decode_level(hm[0, 0], reg[0], cfg.strides[li], cfg)
        if trace is not None:
            trace.setdefault("attn", []).append(attn)
            trace.setdefault("hm", []).append(hm)
            trace.setdefault("reg", []).append(reg)
            trace.setdefault("loc", []).append(loc)
        boxes.append(b)
        scores.append(s)
    boxes, scores = torch.cat(boxes, 0), torch.cat(scores, 0)    # fsod_rpn.py:1109-1110
    keep = proposal_nms_topk(boxes, scores, cfg)
    if trace is not None:
        trace.update(cand_boxes=boxes, cand_scores=scores, keep=keep)
    return boxes[keep], scores[keep]


def detect_image(features: Dict[str, Tensor], protos: Dict[str, Dict[int, Tensor]], sd: Dict[str, Tensor],
                 image_size: Tuple[int, int], cfg: HeadConfig, out_size: Optional[Tuple[int, int]] = None,
                 trace: Optional[dict] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Full head for one image, all classes of the episode (N-way contract of
    SURVEY section 8a: rows Q1..R3 per class, then class-wise NMS over the union).
    ``protos`` has the pkl schema {'p3','p4','p5','rcnn_8','rcnn_4'} -> {cls_id -> Tensor}.
    Returns (pred_boxes [N,4], scores [N], pred_classes [N] int64).  pred_classes is
    the contiguous index of the class in the episode's class list (pkl key order) -
    what the 1-way reference emits (always 0: d2 fast_rcnn.py:169 ``filter_inds[:, 1]``)."""
    cls_ids = list(protos["p3"].keys())
    all_b, all_s, all_c = [], [], []
    raw = [features[l] for l in LEVELS]
    for ci, cid in enumerate(cls_ids):
        tr = {} if trace is not None else None
        pb, ps = propose(features, {l: protos[l][cid] for l in LEVELS}, sd, cfg, tr)
        x = roi_pool(raw, [pb], cfg.pooler_resolution, cfg.strides)
        logits, deltas = relation_head(x, protos["rcnn_8"][cid], sd)
        sc, bx = score_and_decode(logits, deltas, pb, cfg)
        if trace is not None:
            tr.update(proposals=pb, objectness=ps, pooled=x, logits=logits, deltas=deltas, det_scores=sc, det_boxes=bx)
            trace.setdefault("per_class", []).append(tr)
        all_b.append(bx)
        all_s.append(sc)
        all_c.append(torch.full((bx.shape[0],), ci, dtype=torch.int64))
    b, s, c, rows = final_detect(torch.cat(all_b), torch.cat(all_s), torch.cat(all_c), image_size, cfg)
    if trace is not None:
        trace.update(final_rows=rows)
    oh, ow = out_size if out_size is not None else image_size
    b, s, c = postprocess(b, s, c, image_size, oh, ow)
    return b, s, c


__all__ = [n for n in dir() if not n.startswith("_")]
