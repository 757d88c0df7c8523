"""BASELINE.json configs[2] and configs[3] once through the public API (development smoke: shapes, capacities, timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import build_model


def run(name, opts, classes, shots, batch, h, w):
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(bench.ROOT, "configs/fsod/finetune_vovnet.yaml"))
    cfg.merge_from_list(["MODEL.DEVICE", "cuda:0"] + opts)
    model = build_model(cfg).eval()
    model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model.set_prototypes(synth.prototypes(classes, shots, 7))
    imgs = [synth.ore_image(h, w, 1000 + i % 8).pin_memory() for i in range(batch)]
    inputs = [{"image": im} for im in imgs]
    with torch.no_grad():
        for _ in range(2):
            out = model(inputs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = model(inputs)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
    n = [len(o["instances"]) for o in out]
    cls = sorted(set(int(c) for o in out for c in o["instances"].pred_classes.tolist()))
    print(f"{name}: {batch} images {h}x{w}, {len(classes)} classes: {dt * 1e3:.1f} ms/batch = {batch / dt:.0f} img/s; detections/img "
          f"{min(n)}..{max(n)}; classes seen {cls}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


run("configs[2] N-way", ["INPUT.FS.SUPPORT_WAY", 10, "INPUT.FS.SUPPORT_SHOT", 10], list(range(1, 11)), 10, 32, 640, 640)
run("configs[3] high-res", ["MODEL.CENTERNET.PRE_NMS_TOPK_TEST", 2000, "MODEL.CENTERNET.POST_NMS_TOPK_TEST", 2000,
                            "INPUT.FS.SUPPORT_WAY", 2, "INPUT.FS.SUPPORT_SHOT", 5], [1, 2], 5, 16, 800, 1344)
