"""CenterNet2 dense head + proposal generator (host side).

Mirrors the reference interface:
  * ``CenterNetHead`` - CenterNet2/centernet/modeling/dense_heads/centernet_head.py:21-161
    (conv tower + GroupNorm + agn_hm + bbox_pred + Scale).  Row H0 / 8f#1: the 3x3 convolutions
    run on the tensor-core kernel (csrc/conv_tc.cu), GroupNorm on csrc/gn.cu; there is no cuDNN path.
  * ``CenterNet`` - fewx/modeling/fsod/fsod_rpn.py:491-655, 1068-1210, registered in
    PROPOSAL_GENERATOR_REGISTRY under the same name, built as ``cls(cfg, input_shape)``,
    ``forward(images, features_dict, gt_instances) -> (list[Instances], {})``.
    Everything after the convolutions (sigmoid, threshold, top-k, decode, concat, NMS,
    post-NMS top-k) runs in two CUDA kernels over the whole batch with no host sync.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from ..compat import register, resolve, PROPOSAL_GENERATOR_REGISTRY, Boxes, Instances, ShapeSpec
from . import tcconv


class Scale(nn.Module):
    def __init__(self, init_value: float = 1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.FloatTensor([init_value]))

    def forward(self, x):
        return x * self.scale


class CenterNetHead(nn.Module):
    def __init__(self, cfg, input_shape: List[ShapeSpec]):
        super().__init__()
        c = cfg.MODEL.CENTERNET
        if not (c.ONLY_PROPOSAL and c.WITH_AGN_HM):
            raise NotImplementedError("CenterNetHead: only ONLY_PROPOSAL + WITH_AGN_HM (finetune_vovnet.yaml) is built")
        if c.USE_DEFORMABLE or c.NUM_SHARE_CONVS != 0:
            raise NotImplementedError("CenterNetHead: deformable / shared towers are disabled in the reference config")
        in_channels = input_shape[0].channels
        self.only_proposal, self.with_agn_hm = True, True
        tower = []
        for _ in range(c.NUM_BOX_CONVS):
            tower.append(nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1, bias=True))
            if c.NORM == "GN":
                tower.append(nn.GroupNorm(32 if in_channels % 32 == 0 else 25, in_channels))
            elif c.NORM != "":
                raise NotImplementedError(f"CenterNetHead norm {c.NORM}")
            tower.append(nn.ReLU())
        self.cls_tower = nn.Sequential()
        self.bbox_tower = nn.Sequential(*tower)
        self.share_tower = nn.Sequential()
        self.bbox_pred = nn.Conv2d(in_channels, 4, kernel_size=3, stride=1, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in input_shape])
        for m in list(self.bbox_tower.modules()) + [self.bbox_pred]:
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.01)
                nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.bbox_pred.bias, 8.0)
        self.agn_hm = nn.Conv2d(in_channels, 1, kernel_size=3, stride=1, padding=1)
        nn.init.constant_(self.agn_hm.bias, -math.log((1 - c.PRIOR_PROB) / c.PRIOR_PROB))
        nn.init.normal_(self.agn_hm.weight, std=0.01)

    def forward(self, x: Sequence[torch.Tensor], bounds: Optional[Sequence[torch.Tensor]] = None, raw_reg: bool = False):
        """``bounds[l]``: device scalar bounding max|x[l]| when the producer reported it (ops.correlate_levels).
        ``raw_reg``: return bbox_pred's output as it is; the caller applies relu(scale_l * x) (fod_decode_topk does it
        while it reads the map, centernet_head.py:157-160), which saves one elementwise pass per level."""
        clss, bbox_reg, agn_hms = [], [], []
        for l, feature in enumerate(x):
            if not tcconv.supported(self.agn_hm, feature):
                raise ops._lib.FodError("CenterNetHead: fp32 CUDA maps in NHWC memory expected (the head has no cuDNN / CPU path)")
            hm, reg = self._level_tc(feature, bounds[l] if bounds is not None else None)
            clss.append(None)
            agn_hms.append(hm)
            bbox_reg.append(reg if raw_reg else F.relu(self.scales[l](reg)))
        return clss, bbox_reg, agn_hms

    # ---- hot path: tower convolution with GroupNorm statistics in its epilogue -> GroupNorm + ReLU applied to the operand
    # of ONE 1x1 contraction that produces the nine tap products of agn_hm | bbox_pred (45 columns, padded to 48); the nine shifted sums,
    # the biases and Scale + ReLU happen inside fod_decode_topk_taps.  The normalised map, hm and reg never exist.
    def tap_products(self, x: Sequence[torch.Tensor], bounds: Optional[Sequence[torch.Tensor]] = None):
        mods = list(self.bbox_tower)
        if not (len(mods) == 3 and isinstance(mods[0], nn.Conv2d) and isinstance(mods[1], nn.GroupNorm)
                and isinstance(mods[2], nn.ReLU) and mods[1].num_channels % 32 == 0):
            return None
        conv, gn = mods[0], mods[1]
        pk9 = self._packed_taps()
        out = []
        for l, t in enumerate(x):
            if not tcconv.supported(conv, t):
                raise ops._lib.FodError("CenterNetHead: fp32 CUDA maps in NHWC memory expected (the head has no cuDNN / CPU path)")
            n, _, h, w = t.shape
            tiles = ops.conv2d_tiles_per_image(h, w)
            cs = torch.empty((n, tiles, conv.out_channels), dtype=torch.float32, device=t.device)
            cq = torch.empty_like(cs)
            a_t = ops.new_amax(t.device, n)                        # per problem
            b = bounds[l] if bounds is not None else None
            if b is not None and b.numel() == n and n > 1:
                b = b.reshape(1, -1)
            t = tcconv.conv(t, conv, x_amax=b, y_amax=a_t, colsum=cs, colsumsq=cq)
            scale, shift, a_g = ops.group_norm_affine(cs, cq, h * w, gn.num_groups, gn.weight, gn.bias, gn.eps, x_amax=a_t)
            if a_g.numel() == n and n > 1:
                a_g = a_g.reshape(1, -1)
            out.append(ops.conv2d_nhwc(t, pk9, None, 48, 1, x_amax=a_g, a_gate=scale, a_shift=shift, a_relu=True))
        return out

    def _packed_taps(self) -> torch.Tensor:
        """[48, 128, 1, 1]: row tap = filter tap (ky, kx) of agn_hm, row 12 + tap*4 + j = that tap of bbox_pred output j;
        rows 9..11 zero (the layout fod_decode_topk_taps reads)."""
        ws = (self.agn_hm.weight, self.bbox_pred.weight)
        key = tuple((w.data_ptr(), w._version) for w in ws)
        hit = self.__dict__.get("_taps_cache")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                c_in = ws[0].shape[1]
                hm9 = ws[0].permute(0, 2, 3, 1).reshape(9, c_in)                                       # (ky, kx) major
                reg36 = ws[1].permute(2, 3, 0, 1).reshape(36, c_in)                                    # (ky, kx, j) major
                w9 = torch.cat((hm9, hm9.new_zeros((3, c_in)), reg36), 0).reshape(48, c_in, 1, 1).contiguous()
                hit = (key, ops.conv2d_pack(w9.float()))
            self.__dict__["_taps_cache"] = hit
        return hit[1]

    def bias5_host(self) -> List[float]:
        ps = (self.agn_hm.bias, self.bbox_pred.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        hit = self.__dict__.get("_bias5_host")
        if hit is None or hit[0] != key:
            hit = (key, [float(v) for v in torch.cat([p.detach().reshape(-1) for p in ps]).tolist()])
            self.__dict__["_bias5_host"] = hit
        return hit[1]

    def scales_host(self) -> List[float]:
        """The per-level Scale factors as host floats (read back once per weight version)."""
        key = tuple((m.scale.data_ptr(), m.scale._version) for m in self.scales)
        hit = self.__dict__.get("_scales_host")
        if hit is None or hit[0] != key:
            hit = (key, [float(m.scale.detach().reshape(-1)[0].item()) for m in self.scales])
            self.__dict__["_scales_host"] = hit
        return hit[1]

    # conv -> GroupNorm -> ReLU -> conv without writing the normalised map (statistics from the first convolution's
    # epilogue, affine + ReLU on the second one's operand): correct (tests/test_conv_gpu.py) but 0.37 ms per step SLOWER
    # than the separate GroupNorm kernel at batch 64: the narrow output convolution (N = 16) is bound by its operand
    # conversion, which now also waits for two global loads per 4 channels.  Off until those rows are staged in smem.
    FUSE_GROUP_NORM = False

    def _level_tc(self, t: torch.Tensor, bound: Optional[torch.Tensor] = None):
        """Tower and output convolutions on the tensor cores (csrc/conv_tc.cu); agn_hm and bbox_pred read the same
        tower output, so they run as ONE convolution with 1 + 4 (+3 zero) output channels."""
        mods = list(self.bbox_tower)     # ``bound``: device scalar bounding max|t| when the producing kernel reported it
        if (self.FUSE_GROUP_NORM and len(mods) == 3 and isinstance(mods[0], nn.Conv2d) and isinstance(mods[1], nn.GroupNorm)
                and isinstance(mods[2], nn.ReLU) and mods[1].num_channels % 32 == 0 and tcconv.supported(mods[0], t)):
            # conv -> GroupNorm -> ReLU -> conv with the normalised map never written: the tower convolution's epilogue
            # emits per-tile channel sums and sums of squares, a tiny kernel turns them into a per-(problem, channel) scale
            # and shift, and the output convolution applies relu(x * scale + shift) to its input operand
            conv, gn = mods[0], mods[1]
            n, _, h, w = t.shape
            tiles = ops.conv2d_tiles_per_image(h, w)
            cs = torch.empty((n, tiles, conv.out_channels), dtype=torch.float32, device=t.device)
            cq = torch.empty_like(cs)
            a_t = ops.new_amax(t.device)
            t = tcconv.conv(t, conv, x_amax=bound, y_amax=a_t, colsum=cs, colsumsq=cq)
            scale, shift, a_g = ops.group_norm_affine(cs, cq, h * w, gn.num_groups, gn.weight, gn.bias, gn.eps, x_amax=a_t)
            y = tcconv.conv(t, self.agn_hm, extra=self.bbox_pred, x_amax=a_g, a_gate=scale, a_shift=shift, a_relu=True)
            return y[:, 0:1], y[:, 1:5]
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                t, bound = tcconv.conv(t, m, x_amax=None if bound is None else bound.reshape(1, -1) if bound.numel() == t.shape[0] and t.shape[0] > 1 else bound), None
            elif isinstance(m, nn.GroupNorm) and m.num_channels % (4 * m.num_groups) == 0:
                fuse = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)      # GN + ReLU in one pass, in place
                bound = ops.new_amax(t.device, t.shape[0])      # one bound per problem: batch mates do not set a map's scale
                t = ops.group_norm_nhwc(t, m.num_groups, m.weight, m.bias, m.eps, relu=fuse, inplace=True, y_amax=bound)
                i += int(fuse)
            elif isinstance(m, nn.ReLU):
                t = F.relu_(t)          # a bound stays a bound
            else:
                t, bound = m(t).contiguous(memory_format=torch.channels_last), None
            i += 1
        if bound is not None and bound.numel() == t.shape[0] and t.shape[0] > 1:
            bound = bound.reshape(1, -1)                                          # [1, P]: per-image operand scales
        y = tcconv.conv(t, self.agn_hm, extra=self.bbox_pred, x_amax=bound)      # [P, 8, H, W]: hm | l t r b | 0 0 0
        return y[:, 0:1], y[:, 1:5]


@dataclass
class RawProposals:
    """Fixed-capacity batched proposals (device tensors, no host sync)."""
    boxes: torch.Tensor        # [P, roi_cap, 4]
    scores: torch.Tensor       # [P, roi_cap]   objectness (sqrt of the heat-map)
    count: torch.Tensor        # [P] int32
    keep: torch.Tensor         # [P, roi_cap] int64 index into the candidate list
    cand_boxes: torch.Tensor   # [P, cand_cap, 4]
    cand_scores: torch.Tensor  # [P, cand_cap]
    cand_loc: torch.Tensor     # [P, cand_cap] int64
    level_count: torch.Tensor  # [P, L] int32
    cand_count: torch.Tensor   # [P] int32


@register(PROPOSAL_GENERATOR_REGISTRY)
class CenterNet(nn.Module):
    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        c = cfg.MODEL.CENTERNET
        self.in_features = list(c.IN_FEATURES)
        self.strides = list(c.FPN_STRIDES)
        self.score_thresh = c.INFERENCE_TH
        self.pre_nms_topk_test = c.PRE_NMS_TOPK_TEST
        self.post_nms_topk_test = c.POST_NMS_TOPK_TEST
        self.nms_thresh_test = c.NMS_TH_TEST
        self.not_nms = c.NOT_NMS
        self.only_proposal, self.as_proposal, self.with_agn_hm = c.ONLY_PROPOSAL, c.AS_PROPOSAL, c.WITH_AGN_HM
        if c.CENTER_NMS or c.NOT_NMS:
            raise NotImplementedError("CenterNet: CENTER_NMS / NOT_NMS are off in the reference config and not built")
        if len(self.in_features) > 3:
            raise NotImplementedError("CenterNet: at most 3 FPN levels (p3..p5)")
        self.centernet_head = CenterNetHead(cfg, [input_shape[f] for f in self.in_features])
        # extra proposal rows reserved for ties at the post-NMS threshold (fsod_rpn.py:1204 keeps them all)
        self.tie_slack = 64

    FOLD_OUTPUT_CONV = True      # tower -> tap products -> decode (see CenterNetHead.tap_products); False: hm / reg maps

    @property
    def roi_cap(self) -> int:
        return (self.post_nms_topk_test + self.tie_slack + 63) // 64 * 64

    def forward(self, images, features_dict: Dict[str, torch.Tensor], gt_instances=None):
        if self.training:
            raise NotImplementedError("CenterNet: training (targets / losses) is outside the inference hot path")
        features = [features_dict[f] for f in self.in_features]
        status = ops.new_status(features[0].device)
        raw = self.propose_raw(features, status)
        ops.check_status(status)
        n_img = len(images.image_sizes)
        per_image = features[0].shape[0] // max(n_img, 1)
        return self.to_instances(raw, [images.image_sizes[p // per_image] for p in range(features[0].shape[0])]), {}

    @torch.no_grad()
    def propose_raw(self, features: Sequence[torch.Tensor], status: torch.Tensor, roi_cap: Optional[int] = None,
                    bounds: Optional[Sequence[torch.Tensor]] = None) -> RawProposals:
        """features[l]: [P,128,H_l,W_l] correlated maps, one row per (image, class) problem."""
        head = self.centernet_head
        taps = head.tap_products(features, bounds) if self.FOLD_OUTPUT_CONV else None
        if taps is not None:
            boxes, scores, loc, level_count, cand_count = ops.decode_topk_taps(
                taps, head.bias5_host(), self.strides, self.score_thresh, self.pre_nms_topk_test, status,
                reg_scale=head.scales_host())
        else:
            _, reg, hm = head(features, bounds, raw_reg=True)
            boxes, scores, loc, level_count, cand_count = ops.decode_topk(
                hm, reg, self.strides, self.score_thresh, self.pre_nms_topk_test, status, hm_is_logit=True,
                reg_scale=head.scales_host())
        keep, pb, ps, pc = ops.nms_proposals(boxes, scores, cand_count, self.nms_thresh_test, self.post_nms_topk_test,
                                             roi_cap or self.roi_cap, status)
        return RawProposals(pb, ps, pc, keep, boxes, scores, loc, level_count, cand_count)

    @staticmethod
    def to_instances(raw: RawProposals, image_sizes: Sequence[Tuple[int, int]]) -> List[Instances]:
        counts = raw.count.tolist()   # one sync for the whole batch
        out = []
        for p, n in enumerate(counts):
            inst = Instances(tuple(image_sizes[p]))
            inst.scores = raw.scores[p, :n]
            inst.pred_classes = torch.zeros((n,), dtype=torch.int64, device=raw.scores.device)
            inst.proposal_boxes = Boxes(raw.boxes[p, :n])
            inst.objectness_logits = raw.scores[p, :n]
            out.append(inst)
        return out
