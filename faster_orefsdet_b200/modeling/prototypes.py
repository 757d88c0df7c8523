"""Support prototypes: the ``support_feature.pkl`` side channel, its device-resident
form, and the NCCL broadcast.

pkl schema (fsod_cen.py:329,384-389): ``{'p3','p4','p5','rcnn_8','rcnn_4'} -> {cls_id -> CPU fp32 Tensor}``
with shapes [1,128,32,32], [1,128,16,16], [1,128,8,8], [S,128,8,8], [S,128,4,4].

The reference re-reads the pickle and re-uploads every tensor on every forward
(fsod_cen.py:152-153,410-415) and recomputes the seven pooled taps per level per
class per image (:458-460).  Here the episode is reduced once to what the kernels
need - taps [C,7,128] per level, the shot-mean of the pooled support boxes, the
folded per-class bias of the relation head - and stays resident in HBM.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops

LEVELS = ("p3", "p4", "p5")
PKL_KEYS = ("p3", "p4", "p5", "rcnn_8", "rcnn_4")


class MLP(nn.Module):
    """fsod_cen.py:573-582 (dropout is the identity at inference)."""

    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(0.1)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class SM_Block(nn.Module):
    """Support-mixing block of the dense-head prototypes (fsod_cen.py:584-630; printed as
    ``WeightedPermuteMLP`` in log:752-791).  Runs once per episode; stays PyTorch."""

    def __init__(self, dim, seg_dim=8):
        super().__init__()
        self.seg_dim = seg_dim
        self.mlp_h = nn.Linear(dim, dim, bias=False)
        self.mlp_w = nn.Linear(dim, dim, bias=False)
        self.reweighting = MLP(dim, dim // 2, dim * 2)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, H, W, C = x.shape
        S = C // self.seg_dim
        g = x.reshape(B, H, W, self.seg_dim, S)
        h = self.mlp_h(g.permute(0, 3, 2, 1, 4).reshape(B, self.seg_dim, W, H * S))
        h = h.reshape(B, self.seg_dim, W, H, S).permute(0, 3, 2, 1, 4).reshape(B, H, W, C)
        w = self.mlp_w(g.permute(0, 3, 1, 2, 4).reshape(B, self.seg_dim, H, W * S))
        w = w.reshape(B, self.seg_dim, H, W, S).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
        a = (h + w).permute(0, 3, 1, 2).flatten(2).mean(2)
        a = self.reweighting(a).reshape(B, C, 2).permute(2, 0, 1).softmax(0).unsqueeze(2).unsqueeze(2)
        return self.proj_drop(self.proj(w * a[0] + h * a[1]))


@dataclass
class PrototypeBank:
    """Device-resident episode state."""
    class_ids: List[int]
    taps: List[torch.Tensor]        # per level [C,7,128]
    support_mean: torch.Tensor      # [C,128,8,8] shot-mean of rcnn_8 (fsod_roi_heads.py:482)
    bias_cls: torch.Tensor          # [C,128] folded relation-head bias

    @property
    def num_classes(self) -> int:
        return len(self.class_ids)

    @property
    def taps_host(self) -> List[torch.Tensor]:
        """Host copies of the taps (one transfer per episode): fod_correlate_levels takes them as launch parameters."""
        hit = self.__dict__.get("_taps_host")
        if hit is None or hit[0] != tuple((t.data_ptr(), t._version) for t in self.taps):
            hit = (tuple((t.data_ptr(), t._version) for t in self.taps),
                   [t.detach().to("cpu", torch.float32).contiguous() for t in self.taps])
            self.__dict__["_taps_host"] = hit
        return hit[1]

    # ---- one flat fp32 buffer for the NCCL broadcast
    def pack(self) -> torch.Tensor:
        return torch.cat([t.reshape(-1) for t in self.taps] + [self.support_mean.reshape(-1), self.bias_cls.reshape(-1)])

    @staticmethod
    def packed_numel(num_classes: int, num_levels: int = 3) -> int:
        return num_classes * (num_levels * 7 * 128 + 128 * 64 + 128)

    @staticmethod
    def unpack(buf: torch.Tensor, class_ids: Sequence[int], num_levels: int = 3) -> "PrototypeBank":
        C = len(class_ids)
        taps, off = [], 0
        for _ in range(num_levels):
            taps.append(buf[off: off + C * 7 * 128].reshape(C, 7, 128))
            off += C * 7 * 128
        sm = buf[off: off + C * 128 * 64].reshape(C, 128, 8, 8)
        off += C * 128 * 64
        bias = buf[off: off + C * 128].reshape(C, 128)
        return PrototypeBank(list(class_ids), taps, sm, bias)


# ---- binary, mmap-able episode file: what the kernels need, nothing to unpickle
#   bytes 0..63   header: magic "FODB", u32 version, u32 num_classes, u32 num_levels, u64 source mtime_ns, u64 source size,
#                 zero padding
#   then          int64 class_ids[num_classes], padded to a multiple of 64 bytes
#   then          fp32 payload = PrototypeBank.pack() (taps per level, support mean, folded bias): the NCCL broadcast buffer
_FODB_MAGIC = b"FODB"
_FODB_VERSION = 1


def save_bank(bank: PrototypeBank, path: str, source_key=(0, 0)) -> None:
    import struct
    import numpy as np
    ids = np.asarray(bank.class_ids, dtype=np.int64)
    pad = (-ids.nbytes) % 64
    header = _FODB_MAGIC + struct.pack("<IIIQQ", _FODB_VERSION, bank.num_classes, len(bank.taps), int(source_key[0]), int(source_key[1]))
    payload = bank.pack().detach().to("cpu", torch.float32).contiguous().numpy()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(header.ljust(64, b"\0"))
        f.write(ids.tobytes() + b"\0" * pad)
        f.write(payload.tobytes())
    os.replace(tmp, path)      # atomic: readers never see a half-written file


def load_bank(path: str, device, expect_source_key=None) -> Optional[PrototypeBank]:
    """Memory-maps the episode file and uploads its payload with one copy; None if the file is missing, of another
    version, truncated, or (when ``expect_source_key`` = (mtime_ns, size) is given) derived from another pickle."""
    import struct
    import numpy as np
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        head = f.read(64)
    if len(head) < 64 or head[:4] != _FODB_MAGIC:
        return None
    version, C, L, mtime_ns, size = struct.unpack("<IIIQQ", head[4:32])
    if version != _FODB_VERSION or (expect_source_key is not None and (mtime_ns, size) != tuple(int(v) for v in expect_source_key)):
        return None
    ids_bytes = C * 8 + ((-C * 8) % 64)
    n = PrototypeBank.packed_numel(C, L)
    if os.path.getsize(path) != 64 + ids_bytes + 4 * n:
        return None
    ids = np.memmap(path, dtype=np.int64, mode="r", offset=64, shape=(C,))
    payload = np.memmap(path, dtype=np.float32, mode="r", offset=64 + ids_bytes, shape=(n,))
    buf = torch.from_numpy(np.array(payload)).to(device)      # one copy out of the page cache (a memmap is not writable)
    return PrototypeBank.unpack(buf, [int(v) for v in ids], L)


def bank_from_support_dict(support_dict: Dict[str, Dict[int, torch.Tensor]], roi_heads, device) -> PrototypeBank:
    """pkl-schema dict -> PrototypeBank on ``device``.  Taps come from the CUDA kernel (row Q1)."""
    class_ids = list(support_dict["p3"].keys())
    taps = []
    for l in LEVELS:
        proto = torch.cat([support_dict[l][c].to(device=device, dtype=torch.float32) for c in class_ids], 0)
        taps.append(ops.support_taps(proto))
    sm = torch.cat([support_dict["rcnn_8"][c].to(device=device, dtype=torch.float32).mean(0, True) for c in class_ids], 0)
    return PrototypeBank(class_ids, taps, sm.contiguous(), roi_heads.class_bias(sm))


def broadcast_bank(bank: Optional[PrototypeBank], device, src: int = 0) -> PrototypeBank:
    """Rank ``src`` holds the bank; every other rank receives it with one NCCL broadcast of the
    class-id list and one of the packed fp32 buffer (replaces every rank reading the pickle
    on every forward)."""
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = bank.num_classes
    dist.broadcast(meta, src)
    C = int(meta.item())
    ids = torch.zeros(C, dtype=torch.int64, device=device)
    if rank == src:
        ids.copy_(torch.tensor(bank.class_ids, dtype=torch.int64))
    dist.broadcast(ids, src)
    buf = bank.pack().to(device) if rank == src else torch.empty(PrototypeBank.packed_numel(C), dtype=torch.float32, device=device)
    dist.broadcast(buf, src)
    return PrototypeBank.unpack(buf, ids.tolist())


class SupportCache:
    """``./support_dir/support_feature.pkl`` reader, cached on (mtime, size) instead of being
    re-read per forward."""

    def __init__(self, path: str = os.path.join("support_dir", "support_feature.pkl")):
        self.path = path
        self._key = None
        self._dict = None

    def exists(self) -> bool:
        return os.path.exists(self.path)

    def load(self) -> Dict[str, Dict[int, torch.Tensor]]:
        st = os.stat(self.path)
        key = (os.path.abspath(self.path), st.st_mtime_ns, st.st_size)
        if key != self._key:
            with open(self.path, "rb") as f:
                d = pickle.load(f, encoding="latin1")
            for k in PKL_KEYS:
                if k not in d:
                    raise KeyError(f"{self.path}: missing '{k}' (expected keys {PKL_KEYS})")
            self._key, self._dict = key, d
        return self._dict

    @property
    def key(self):
        return self._key

    def stat_key(self):
        """(mtime_ns, size) of the pickle without reading it."""
        st = os.stat(self.path)
        return (st.st_mtime_ns, st.st_size)

    @property
    def binary_path(self) -> str:
        return os.path.splitext(self.path)[0] + ".fodb"
