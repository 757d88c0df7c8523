"""Host glue for the tensor-core convolution (``fod_conv2d_nhwc``, csrc/conv_tc.cu).

A ``torch.nn.Conv2d`` (optionally followed by a frozen BatchNorm that is folded into it) is
run on CUDA tensors through the C-ABI kernel: fp16-split operands on tcgen05 = fp32 accuracy, bias and
ReLU fused, NHWC in and out, input and output allowed to be channel slices of wider NHWC
buffers (so an OSA concat is written in place).  There is no switch back to cuDNN.  CPU tensors take
PyTorch's own convolution: the backbone module is also what the CPU baseline of bench.py runs on the
host cores.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops

_cache: "weakref.WeakKeyDictionary[nn.Module, tuple]" = weakref.WeakKeyDictionary()


def supported(conv: nn.Conv2d, x: torch.Tensor) -> bool:
    k = conv.kernel_size[0]
    return (x.is_cuda and x.dtype == torch.float32 and conv.kernel_size in ((1, 1), (3, 3))
            and (conv.stride == (1, 1) or (conv.stride == (2, 2) and k == 3))
            and conv.padding == (k // 2, k // 2) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.in_channels % 4 == 0)


def folded(conv: nn.Conv2d, norm: Optional[nn.Module] = None):
    """(weight, bias) of conv followed by a frozen BatchNorm (d2!/layers/batch_norm.py FrozenBatchNorm2d)."""
    w, b = conv.weight, conv.bias
    if norm is not None:
        scale = norm.weight * (norm.running_var + norm.eps).rsqrt()
        shift = norm.bias - norm.running_mean * scale
        w = w * scale.reshape(-1, 1, 1, 1)
        b = shift if b is None else b * scale + shift
    return w, b


def _versions(*ts):
    return tuple((t.data_ptr(), t._version) for t in ts if t is not None)


def packed(conv: nn.Conv2d, norm: Optional[nn.Module] = None, extra: Optional[nn.Conv2d] = None):
    """Packed tf32 hi/lo weight planes + bias of ``conv`` (BN folded); ``extra`` stacks a second convolution's
    output channels behind the first (agn_hm + bbox_pred share their input).  Output channels are padded to a
    multiple of 4 with zero filters.  Cached per module until a parameter changes."""
    key = _versions(conv.weight, conv.bias, *(list(norm.buffers()) if norm is not None else []),
                    *((extra.weight, extra.bias) if extra is not None else ()))
    hit = _cache.get(conv)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2], hit[3]
    with torch.no_grad():
        w, b = folded(conv, norm)
        if extra is not None:
            w = torch.cat((w, extra.weight), 0)
            zb = lambda m: m.bias if m.bias is not None else torch.zeros(m.out_channels, device=w.device)
            b = torch.cat((b if b is not None else zb(conv), zb(extra)), 0)
        cout = w.shape[0]
        pad = (-cout) % 4
        if pad:
            w = torch.cat((w, w.new_zeros((pad,) + tuple(w.shape[1:]))), 0)
            if b is not None:
                b = torch.cat((b, b.new_zeros(pad)), 0)
        pk = ops.conv2d_pack(w.float().contiguous())
        b = None if b is None else b.detach().float().contiguous()
    _cache[conv] = (key, pk, b, cout + pad)
    return pk, b, cout + pad


def conv(x: torch.Tensor, m: nn.Conv2d, norm: Optional[nn.Module] = None, relu: bool = False,
         out: Optional[torch.Tensor] = None, extra: Optional[nn.Conv2d] = None, x_amax: Optional[torch.Tensor] = None,
         y_amax: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         residual_upsample2: bool = False, a_gate: Optional[torch.Tensor] = None,
         colsum: Optional[torch.Tensor] = None, a_shift: Optional[torch.Tensor] = None, a_relu: bool = False,
         colsumsq: Optional[torch.Tensor] = None, x_presplit: bool = False) -> torch.Tensor:
    """conv (+ folded frozen BN) (+ ReLU) of an NHWC CUDA view; returns [N, Cout_padded_to_4, H, W] (NHWC memory).
    ``x_amax`` / ``y_amax``: device scalars bounding max|x| (computed when omitted) / receiving max|y| (ops.conv2d_nhwc)."""
    pk, b, cout = packed(m, norm, extra)
    return ops.conv2d_nhwc(x, pk, b, cout, m.kernel_size[0], relu, out=out, stride=m.stride[0], x_amax=x_amax, y_amax=y_amax,
                           residual=residual, residual_upsample2=residual_upsample2, a_gate=a_gate, colsum=colsum,
                           a_shift=a_shift, a_relu=a_relu, colsumsq=colsumsq, x_presplit=x_presplit)


_bound_cache: "weakref.WeakKeyDictionary[nn.Module, tuple]" = weakref.WeakKeyDictionary()


def bound_consts(conv: nn.Conv2d, norm: Optional[nn.Module] = None):
    """(l1, beta) with  max|relu(conv(x))| <= l1 * max|x| + beta  for any x: the largest absolute row sum of the folded
    weight and the largest |bias| (+ 0.1 % for the rounding of the sums).  Cached per module until a parameter changes."""
    key = _versions(conv.weight, conv.bias, *(list(norm.buffers()) if norm is not None else []))
    hit = _bound_cache.get(conv)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            w, b = folded(conv, norm)
            l1 = float(w.double().abs().sum((1, 2, 3)).max()) * 1.001
            beta = float(b.double().abs().max()) * 1.001 if b is not None else 0.0
        hit = (key, l1, beta)
        _bound_cache[conv] = hit
    return hit[1], hit[2]


def conv_reference(x: torch.Tensor, m: nn.Conv2d, norm: Optional[nn.Module] = None, relu: bool = False) -> torch.Tensor:
    """The same operation with PyTorch's convolution (CPU tensors, unsupported shapes)."""
    w, b = folded(m, norm)
    y = F.conv2d(x, w, b, m.stride, m.padding, m.dilation, m.groups)
    return F.relu_(y) if relu else y
