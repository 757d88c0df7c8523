"""Where a FsodRCNN step goes (development tool): torch.profiler kernel table of model(batched_inputs), batch 16 x 640x640."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.compat import META_ARCH_REGISTRY
from faster_orefsdet_b200.config import get_cfg
import faster_orefsdet_b200.modeling  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
cfg = get_cfg()
cfg.merge_from_file(os.path.join(ROOT, "configs/fsod/Base-FSOD-C4.yaml"))
cfg.merge_from_list(["MODEL.DEVICE", "cuda:0"])
m = META_ARCH_REGISTRY.get("FsodRCNN")(cfg).to(dev).eval()
m.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
sup = {"res4_avg": {}, "res5_avg": {}}
for j, c in enumerate((3, 9)):
    sup["res4_avg"][c] = synth.tensor((1, 1024, 14, 14), 341 + 2 * j, 0.0, 1.2)
    sup["res5_avg"][c] = synth.tensor((1, 2048, 7, 7), 342 + 2 * j, 0.0, 1.0)
m.set_prototypes(sup)
batch = [{"image": synth.ore_image(640, 640, 7000 + i).to(dev)} for i in range(16)]
with torch.no_grad():
    for _ in range(3):
        m(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        m(batch)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
