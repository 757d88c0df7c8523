"""Algebraic folding of the relation head's linear chain (done once per weight load
/ once per episode, in float64, result stored fp32).

Reference (fewx/modeling/fsod/fsod_roi_heads.py:509-511; layers created in the
vendored detectron2 roi_heads.py:585-592, box_head.py:66-74):

    A = conv3(cat(x, s)) + cat(conv1(x), conv2(s))          # 1x1 convs, x,s: [128,8,8]
    f = relu(fc1(flatten(A)))                                 # 8192 -> 128

There is no non-linearity between the 1x1 convs and fc1, so with
Wx[o,c] = W3[o,c] + (W1[o,c] if o < 64 else 0):

    f = relu( sum_{c,bin} Wfold[j, bin, c] * x[c, bin] + bias_cls[j] )
    Wfold[j, bin, c] = sum_o Wfc[j, o, bin] * Wx[o, c]                  (class independent)
    bias_cls[j]      = sum_{o,bin} Wfc[j, o, bin] * K[o, bin] + bfc[j]   (per support class)
    K[o, bin]        = sum_c W3[o,128+c] s[c,bin] + b3[o] + (b1[o] if o<64 else sum_c W2[o-64,c] s[c,bin] + b2[o-64])

The folded K index is bin*128 + c, matching the NHWC pooled layout written by
fod_roi_align ([roi][bin][channel]).  The reassociation changes results at the
1e-7 relative level (measured in tests/test_fold.py), far inside the 1e-4 budget.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor


def fold_relation_weights(sd: Dict[str, Tensor], prefix: str = "roi_heads.") -> Tuple[Tensor, Tensor, Tensor]:
    """-> (w_fold [128, 8192], w_out [6, 128], b_out [6]) fp32."""
    d = torch.float64
    w3 = sd[prefix + "conv3.weight"].to(d).reshape(128, 256)
    w1 = sd[prefix + "conv1.weight"].to(d).reshape(64, 128)
    wfc = sd[prefix + "box_head.0.fc1.weight"].to(d).reshape(128, 128, 64)  # [j, o, bin]
    wx = w3[:, :128].clone()
    wx[:64] += w1
    w_fold = torch.einsum("job,oc->jbc", wfc, wx).reshape(128, 64 * 128)
    w_out = torch.cat((sd[prefix + "box_predictor.0.cls_score.weight"], sd[prefix + "box_predictor.0.bbox_pred.weight"]), 0)
    b_out = torch.cat((sd[prefix + "box_predictor.0.cls_score.bias"], sd[prefix + "box_predictor.0.bbox_pred.bias"]), 0)
    return w_fold.float().contiguous(), w_out.float().contiguous(), b_out.float().contiguous()


def fold_class_bias(sd: Dict[str, Tensor], support_mean: Tensor, prefix: str = "roi_heads.") -> Tensor:
    """support_mean [C,128,8,8] (shot-mean of the pooled support boxes, fsod_roi_heads.py:482)
    -> bias_cls [C,128] fp32."""
    d = torch.float64
    C = support_mean.shape[0]
    s = support_mean.to(d).reshape(C, 128, 64)
    w3s = sd[prefix + "conv3.weight"].to(d).reshape(128, 256)[:, 128:]
    w2 = sd[prefix + "conv2.weight"].to(d).reshape(64, 128)
    k = torch.einsum("oc,ncb->nob", w3s, s) + sd[prefix + "conv3.bias"].to(d)[None, :, None]
    k[:, :64] += sd[prefix + "conv1.bias"].to(d)[None, :, None]
    k[:, 64:] += torch.einsum("oc,ncb->nob", w2, s) + sd[prefix + "conv2.bias"].to(d)[None, :, None]
    wfc = sd[prefix + "box_head.0.fc1.weight"].to(d).reshape(128, 128, 64)
    bias = torch.einsum("job,nob->nj", wfc, k) + sd[prefix + "box_head.0.fc1.bias"].to(d)[None]
    return bias.float().contiguous()
