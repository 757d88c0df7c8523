"""Where a step goes: backbone vs CenterNetHead towers vs CUDA head (CUDA events; development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.modeling import build_model
B = 64
model = build_model(bench._cfg("cuda:0")).eval()
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
model.load_state_dict(synth.state_dict(shapes))
model.set_prototypes(synth.prototypes([1], bench.SHOTS, 7))
x = torch.stack(bench._images(B, 1000)).cuda()
x = ((x.float() - model.pixel_mean) / model.pixel_std).contiguous(memory_format=torch.channels_last)


def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    feats = model.backbone(x)
    raw = [feats[f] for f in model.in_features]
    attn = __import__("faster_orefsdet_b200").ops.correlate_levels(raw, model._bank.taps, model.conv3.weight, model.conv3.bias)
    print("backbone ms", timeit(lambda: model.backbone(x)))
    print("towers ms", timeit(lambda: model.proposal_generator.centernet_head(attn)))
    sizes = [(640, 640)] * B
    print("head (incl towers) ms", timeit(lambda: model.head(feats, sizes, sizes)))
    for l, a in enumerate(attn):
        t = model.proposal_generator.centernet_head.bbox_tower
        print(f" level {l} tower conv+GN+relu ms", timeit(lambda: t(a)))
