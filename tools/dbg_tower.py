"""Per-stage error of the CenterNetHead tower (tensor-core path vs cuDNN fp32 path) against float64 on the CPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from faster_orefsdet_b200 import synth, ops
from faster_orefsdet_b200.modeling import tcconv
from tests.test_model_gpu import _model
model = _model()
model.set_prototypes(synth.prototypes([1], 5, 7))
head = model.proposal_generator.centernet_head
feats = synth.features(2, 256, 320, 91)
raw = [feats[f].cuda().contiguous(memory_format=torch.channels_last) for f in model.in_features]
attn = ops.correlate_levels(raw, model._bank.taps, model.conv3.weight, model.conv3.bias)
conv, gn = head.bbox_tower[0], head.bbox_tower[1]


def err(a, b):
    b = b.double()
    return float((a.double().cpu() - b).abs().max() / b.abs().max())


for l, a in enumerate(attn):
    a64 = a.double().cpu()
    t64 = F.conv2d(a64, conv.weight.double().cpu(), conv.bias.double().cpu(), padding=1)
    g64 = F.relu(F.group_norm(t64, 32, gn.weight.double().cpu(), gn.bias.double().cpu(), gn.eps))
    w5 = torch.cat((head.agn_hm.weight, head.bbox_pred.weight)).double().cpu()
    b5 = torch.cat((head.agn_hm.bias, head.bbox_pred.bias)).double().cpu()
    o64 = F.conv2d(g64, w5, b5, padding=1)
    for name, on in (("tc", True), ("cudnn", False)):
        tcconv.ENABLED = on
        if on:
            t = tcconv.conv(a, conv)
            g = ops.group_norm_nhwc(t.clone(), 32, gn.weight, gn.bias, gn.eps, relu=True)
            o = tcconv.conv(g, head.agn_hm, extra=head.bbox_pred)[:, :5]
        else:
            t = conv(a)
            g = F.relu(gn(t))
            o = torch.cat((head.agn_hm(g), head.bbox_pred(g)), 1)
        # error of each stage given ITS OWN exact input is what the kernel is responsible for
        t_in = F.conv2d(a64, conv.weight.double().cpu(), conv.bias.double().cpu(), padding=1)
        g_own = F.relu(F.group_norm(t.double().cpu(), 32, gn.weight.double().cpu(), gn.bias.double().cpu(), gn.eps))
        o_own = F.conv2d(g.double().cpu(), w5, b5, padding=1)
        print(f"level {l} {name:5s}: tower conv {err(t, t_in):.2e}  GN(own input) {err(g, g_own):.2e}  out conv(own input) {err(o, o_own):.2e}"
              f"  | end-to-end: GN {err(g, g64):.2e} out {err(o, o64):.2e}  hm-only {err(o[:, :1], o64[:, :1]):.2e}"
              f"  mean signed rel of tower {float(((t.double().cpu() - t_in) / t_in.abs().clamp_min(1e-3 * t_in.abs().max())).mean()):.2e}")
