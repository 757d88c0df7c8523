"""A/B of the two ROIAlign kernels (tile-stationary | one CTA per ROI) on the detector's own FPN maps and proposals
(bench configuration: batch 64 x 640x640, 1-way 25-shot), outputs compared bit for bit.  `python tools/bench_roi.py [B] [once]`:
with `once` the captured call is replayed three times per kernel and nothing is timed (the ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import build_model
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
once = len(sys.argv) > 2
cfg = get_cfg()
cfg.merge_from_file(os.path.join(ROOT, "configs/fsod/finetune_vovnet.yaml"))
cfg.merge_from_list(["MODEL.DEVICE", "cuda", "INPUT.FS.SUPPORT_SHOT", 25])
model = build_model(cfg).eval()
model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
model.set_prototypes(synth.prototypes([1], 25, 7))
model.USE_CUDA_GRAPH = False
base = [synth.ore_image(640, 640, 1000 + i) for i in range(8)]
x = torch.stack([torch.roll(base[i % 8], shifts=(7 * (i // 8), 13 * (i // 8)), dims=(1, 2)) for i in range(B)]).cuda()
sizes = [(640, 640)] * B
captured = {}
real = ops.roi_align


def spy(feats, strides, rois, roi_count, C, res, **kw):
    if res == 8 and kw.get("tiled"):
        captured["args"] = ([f.clone() for f in feats], strides, rois.clone(), roi_count.clone(), C, res)
    return real(feats, strides, rois, roi_count, C, res, **kw)


ops.roi_align = spy
with torch.no_grad():
    model.detect_from_uint8(x, sizes, sizes)
ops.roi_align = real
torch.cuda.synchronize()
feats, strides, rois, counts, C, res = captured["args"]
print("proposals per image: mean %.1f, min %d, max %d; cap %d" % (counts.float().mean().item(), counts.min().item(), counts.max().item(), rois.shape[1]))
P, cap = rois.shape[0], rois.shape[1]
outs = {}
for name, per_roi in (("tile", False), ("per_roi", True)):
    out = torch.zeros((P, (cap + 127) // 128, 256, 128, 32), device="cuda")
    for _ in range(3):
        real(feats, strides, rois, counts, C, res, out=out, tiled=True, per_roi=per_roi)
    torch.cuda.synchronize()
    outs[name] = out
    if once:
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(10):
            real(feats, strides, rois, counts, C, res, out=out, tiled=True, per_roi=per_roi)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"roi_align[{name}] {best * 1e3:.1f} us per call (all launches of the call, best of 5 x 10)")
print("bit-identical:", torch.equal(outs["tile"], outs["per_roi"]))
