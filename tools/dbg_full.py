"""Replays tests/test_model_gpu.py::test_full_detector_forward_with_pkl_side_channel and prints both detection lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import synth
from oracle import head_oracle as O
from tests.test_model_gpu import _model
torch.set_printoptions(precision=6, linewidth=200, sci_mode=False)
protos = synth.prototypes([1], 5, 7)
model = _model()
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
model.load_state_dict(synth.state_dict(shapes), strict=False)
model.set_prototypes(protos)
imgs = [synth.ore_image(256, 320, 1000), synth.ore_image(224, 300, 1001)]
inputs = [{"image": imgs[0], "height": 300, "width": 375}, {"image": imgs[1]}]
out = model(inputs)
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
images = model.preprocess_image(inputs)
with torch.no_grad():
    feats = {k: v.cpu().contiguous() for k, v in model.backbone(images.tensor).items()}
for b, inp in enumerate(inputs):
    size = images.image_sizes[b]
    outsz = (inp.get("height", size[0]), inp.get("width", size[1]))
    rb, rs, rc = O.detect_image({k: v[b:b + 1] for k, v in feats.items()}, protos, sd, size, O.HeadConfig(), outsz)
    inst = out[b]["instances"]
    gb, gs = inst.pred_boxes.tensor.cpu(), inst.scores.cpu()
    print("image", b, "ref", rb.shape[0], "got", gb.shape[0])
    for i in range(max(rb.shape[0], gb.shape[0])):
        r = (rb[i].tolist(), float(rs[i])) if i < rb.shape[0] else None
        g = (gb[i].tolist(), float(gs[i])) if i < gb.shape[0] else None
        print(i, "REF", r and ["%.4f" % v for v in r[0]], r and "%.6e" % r[1], "GOT", g and ["%.4f" % v for v in g[0]], g and "%.6e" % g[1])
