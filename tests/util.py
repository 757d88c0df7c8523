"""Shared helpers for the tests: golden fixtures, synthetic weights, comparisons."""
import os

import numpy as np
import torch

from faster_orefsdet_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def head_param_shapes():
    shapes = {}
    with open(os.path.join(GOLDEN, "head_param_shapes.txt")) as f:
        for line in f:
            parts = line.split()
            shapes[parts[0]] = tuple(int(x) for x in parts[1:])
    return shapes


def head_state_dict():
    """The synthetic head weights the golden vectors were produced with."""
    return synth.state_dict(head_param_shapes())


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def assert_close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(b).double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.numel() == 0:
        return
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{a.numel()} elements off; max abs err "
                           f"{float(err.max()):.3e}, max rel {float((err / (b.abs() + 1e-12)).max()):.3e}")
