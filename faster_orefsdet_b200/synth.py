"""Deterministic synthetic tensors shared by the golden-vector generator, the
tests, ``smoke()`` and ``bench.py``.

Everything is produced by an integer hash (splitmix64) of (seed, element index),
so the same values come out on every machine and every library version; golden
fixtures therefore store only OUTPUTS plus the seeds/shapes of their inputs.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Sequence, Tuple

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform01(shape: Sequence[int], seed: int) -> np.ndarray:
    """float64 array of multiples of 2^-24 in [0, 1)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + (np.uint64(seed) * np.uint64(0x100000001B3) & _M64)
        h = _splitmix64(idx)
    return ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).reshape(shape)


def tensor(shape: Sequence[int], seed: int, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    u = uniform01(tuple(shape), seed)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32))


def name_seed(name: str, base: int = 0) -> int:
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ base) & 0x7FFFFFFF


def state_dict(shapes: Dict[str, Tuple[int, ...]], base_seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic weights for every key.  Magnitudes follow the reference's
    initialisers (centernet_head.py:89-121, d2 fast_rcnn.py:384-387, PyTorch
    defaults) closely enough that every stage of the head does real work:
    all three FPN levels receive proposals, NMS suppresses most candidates, the
    relation logits are not saturated."""
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        seed = name_seed(name, base_seed)
        leaf = name.rsplit(".", 1)[-1]
        if "num_batches_tracked" in name:
            out[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        if leaf == "running_var":
            out[name] = tensor(shape, seed, 0.5, 1.5)
        elif leaf == "running_mean":
            out[name] = tensor(shape, seed, -0.2, 0.2)
        elif leaf == "scale":
            out[name] = tensor(shape, seed, 0.9, 1.1)
        elif name.endswith("centernet_head.bbox_pred.bias"):
            out[name] = tensor(shape, seed, 6.0, 10.0)
        elif name.endswith("centernet_head.agn_hm.bias"):
            out[name] = tensor(shape, seed, -4.7, -4.5)
        elif name.endswith("centernet_head.agn_hm.weight"):
            out[name] = tensor(shape, seed, -0.06, 0.06)
        elif name.endswith("centernet_head.bbox_pred.weight"):
            out[name] = tensor(shape, seed, -0.12, 0.12)
        elif any(t in name for t in ("cls_score_cor.weight", "cls_score_fc.weight", "cls_score_pr.weight")):
            out[name] = tensor(shape, seed, -0.0004, 0.0004)     # FsodFastRCNNOutputLayers: logits stay unsaturated
        elif "bbox_pred_cor.weight" in name:
            out[name] = tensor(shape, seed, -0.0004, 0.0004)
        elif "cls_score.weight" in name:
            out[name] = tensor(shape, seed, -0.25, 0.25)
        elif "box_predictor.0.bbox_pred.weight" in name:
            out[name] = tensor(shape, seed, -0.1, 0.1)
        elif leaf == "bias":
            out[name] = tensor(shape, seed, -0.1, 0.1)
        elif leaf == "weight" and len(shape) == 1:        # norm scales
            out[name] = tensor(shape, seed, 0.8, 1.2)
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
            b = math.sqrt(3.0 / fan_in)
            out[name] = tensor(shape, seed, -b, b)
        else:
            out[name] = tensor(shape, seed, -0.1, 0.1)
    return out


def features(batch: int, height: int, width: int, seed: int, channels: int = 128,
             strides: Iterable[int] = (8, 16, 32)) -> Dict[str, torch.Tensor]:
    """Stand-in for the backbone's FPN output on a (height x width) padded image."""
    out = {}
    for i, s in enumerate(strides):
        out[f"p{3 + i}"] = tensor((batch, channels, height // s, width // s), seed * 16 + i, -1.5, 1.5)
    return out


def prototypes(class_ids: Sequence[int], shots: int, seed: int, channels: int = 128) -> Dict[str, Dict[int, torch.Tensor]]:
    """A ``support_feature.pkl``-shaped dict (fsod_cen.py:329,384-389) of synthetic entries."""
    d = {"p3": {}, "p4": {}, "p5": {}, "rcnn_8": {}, "rcnn_4": {}}
    for j, c in enumerate(class_ids):
        s = seed * 131 + j * 7
        d["p3"][c] = tensor((1, channels, 32, 32), s + 0, -0.6, 0.8)
        d["p4"][c] = tensor((1, channels, 16, 16), s + 1, -0.6, 0.8)
        d["p5"][c] = tensor((1, channels, 8, 8), s + 2, -0.6, 0.8)
        d["rcnn_8"][c] = tensor((shots, channels, 8, 8), s + 3, -1.0, 1.0)
        d["rcnn_4"][c] = tensor((shots, channels, 4, 4), s + 4, -1.0, 1.0)
    return d


def ore_image(height: int, width: int, seed: int) -> torch.Tensor:
    """uint8 [3,H,W] grey-replicated ore-like image: dark background, 15-25
    bright textured ellipses."""
    u = uniform01((64,), seed * 977 + 5)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    img = 45.0 + 30.0 * (uniform01((height, width), seed * 977 + 1) - 0.5)
    n = 15 + int(u[0] * 11)
    pu = uniform01((n, 6), seed * 977 + 2)
    tex = 24.0 * (uniform01((height, width), seed * 977 + 3) - 0.5)
    for k in range(n):
        cx, cy = pu[k, 0] * width, pu[k, 1] * height
        rx = (15 + 45 * pu[k, 2]) * width / 640.0
        ry = (15 + 45 * pu[k, 3]) * height / 640.0
        m = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
        img[m] = 80.0 + 60.0 * pu[k, 4] + tex[m]
    g = torch.from_numpy(np.clip(img, 0, 255).astype(np.uint8))
    return g.unsqueeze(0).expand(3, -1, -1).contiguous()
