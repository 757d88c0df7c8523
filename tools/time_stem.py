"""stem_1 / stem_2 / stem_3 inside the detector's own eager step, by CUDA events around each operator call, for the
tensor-core and the FMA stem_1 (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import build_model
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = get_cfg()
cfg.merge_from_file(os.path.join(ROOT, "configs/fsod/finetune_vovnet.yaml"))
cfg.merge_from_list(["MODEL.DEVICE", "cuda", "INPUT.FS.SUPPORT_SHOT", 25])
model = build_model(cfg).eval()
model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
model.set_prototypes(synth.prototypes([1], 25, 7))
model.USE_CUDA_GRAPH = False
B = 64
xs = [torch.stack([synth.ore_image(640, 640, 1000 + 7 * s + i % 8) for i in range(B)]).cuda() for s in range(2)]
sizes = [(640, 640)] * B
rec = {}
for name in ("stem1_u8_tc", "stem1_u8", "conv2d_nhwc"):
    fn = getattr(ops, name)
    def timed(*a, __fn=fn, __n=name, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = __fn(*a, **k)
        e.record()
        rec.setdefault(__n, []).append((s, e))
        return r
    setattr(ops, name, timed)
vov = model.backbone.bottom_up
for tc in (True, False, True, False):
    vov.STEM1_TENSOR_CORES = tc
    with torch.no_grad():
        for k in range(3):
            model.detect_from_uint8(xs[k % 2], sizes, sizes)
        torch.cuda.synchronize()
        rec.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(5):
            model.detect_from_uint8(xs[k % 2], sizes, sizes)
        t1.record()
        torch.cuda.synchronize()
    key = "stem1_u8_tc" if tc else "stem1_u8"
    st = sum(s.elapsed_time(e) for s, e in rec[key]) / len(rec[key])
    convs = rec["conv2d_nhwc"]
    per_step = len(convs) // 5
    c2 = sum(convs[i * per_step][0].elapsed_time(convs[i * per_step][1]) for i in range(5)) / 5
    c3 = sum(convs[i * per_step + 1][0].elapsed_time(convs[i * per_step + 1][1]) for i in range(5)) / 5
    print(f"tensor cores {tc}: eager step {t0.elapsed_time(t1) / 5:.3f} ms; stem_1 {st:.3f} ms, stem_2 {c2:.3f} ms, stem_3 {c3:.3f} ms")
