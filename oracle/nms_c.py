"""ctypes binding of oracle/nms_ref.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libnms_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(_PATH)
        _lib.fod_oracle_batched_nms.restype = ctypes.c_int64
        _lib.fod_oracle_batched_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_double, ctypes.c_void_p]
    return _lib


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs, thr: float) -> torch.Tensor:
    b = np.ascontiguousarray(boxes.numpy(), dtype=np.float32)
    s = np.ascontiguousarray(scores.numpy(), dtype=np.float32)
    n = b.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    ip = None
    if idxs is not None:
        i64 = np.ascontiguousarray(idxs.numpy(), dtype=np.int64)
        ip = i64.ctypes.data
    k = _load().fod_oracle_batched_nms(b.ctypes.data, s.ctypes.data, ip, n, float(thr), keep.ctypes.data)
    return torch.from_numpy(keep[:k].copy())
