"""Error of the feature extractor paths against a float64 CPU evaluation of the same module (development tool)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.modeling import tcconv
from tests.test_model_gpu import _model
model = _model()
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
model.load_state_dict(synth.state_dict(shapes), strict=False)
imgs = [synth.ore_image(128, 160, 3000 + i) for i in range(5)]
inputs = [{"image": im} for im in imgs]
ref_model = copy.deepcopy(model.backbone).cpu().double()
x = model.preprocess_image(inputs).tensor
with torch.no_grad():
    ref = ref_model(x.cpu().double())
    a, _ = model._features_pipelined(inputs)
    b = model.backbone(x)
    tcconv.ENABLED = False
    c = model.backbone(x)
    tcconv.ENABLED = True
for k in ref:
    m = float(ref[k].abs().max())
    e = lambda t: float((t.double().cpu() - ref[k]).abs().max()) / m
    print(f"{k}: max|ref| {m:.3e}   u8+FFMA-stem path {e(a[k]):.2e}   float+im2col path {e(b[k]):.2e}   cuDNN fp32 {e(c[k]):.2e}   (relative to max)")
