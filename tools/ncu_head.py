"""One eager pass of the detector at the bench configuration (batch 64 x 640x640, 1-way 25-shot) for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'correlate_tc|decode_topk|nms_proposals|roi_|relation_tc|final_detect' \\
        -s <launches of the 2 warm-up passes> -c 8 -o gpurun_out/r2_head python tools/ncu_head.py
(development tool; prints the kernel launch order of one pass so that -s can be chosen)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import build_model
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
with torch.no_grad():
    for _ in range(passes):
        out = model.detect_from_uint8(x, sizes, sizes)
torch.cuda.synchronize()
print("ok", int(out[3].sum()))
