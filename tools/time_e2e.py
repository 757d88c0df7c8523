"""Where the end-to-end call spends host time (development tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.modeling import build_model, roi_heads
model = build_model(bench._cfg("cuda:0")).eval()
model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
model.set_prototypes(synth.prototypes([1], 25, 7))
imgs = [im.pin_memory() for im in bench._images(64, 1000)]
inputs = [{"image": im} for im in imgs]
T = {}


def timed(obj, name):
    fn = getattr(obj, name)

    def w(*a, **k):
        torch.cuda.synchronize() if name == "pack_instances" else None
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T.setdefault(name, []).append(time.perf_counter() - t0)
        return r
    setattr(obj, name, w)


with torch.no_grad():
    for _ in range(3):
        model(inputs)
    for n in ("_stage_uint8", "detect_from_uint8", "_graph_key", "_stem_from_uint8", "_graph_replay", "_head_finish"):
        timed(model, n)
    import faster_orefsdet_b200.modeling.fsod_cen as fc
    timed(fc, "pack_instances")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = model(inputs)
        for r in out:
            r["instances"].to("cpu")
    torch.cuda.synchronize()
    print("e2e ms/batch", (time.perf_counter() - t0) / 10 * 1e3)
for k, v in T.items():
    print(f"{k:20s} {sum(v) / len(v) * 1e3:7.3f} ms (host wall, includes waiting on the device where it syncs)")
