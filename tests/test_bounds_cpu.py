"""Host side of the split hand-off between convolutions: the bounds that fix the fp16 scale of a
producer's output BEFORE it runs must hold for every input - checked here on the CPU against PyTorch's convolutions."""
import torch
from torch.nn import functional as F

from faster_orefsdet_b200 import synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import META_ARCH_REGISTRY, tcconv


def _vov():
    cfg = get_cfg()
    import os
    cfg.merge_from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs/fsod/finetune_vovnet.yaml"))
    model = META_ARCH_REGISTRY.get("CenterNet2Detector")(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    return model, model.backbone.bottom_up


def test_layer_bound_constants_hold_for_any_input():
    """max|relu(conv(x))| <= l1 * max|x| + beta, with equality approached by the sign pattern of the heaviest filter."""
    _, vov = _vov()
    mod = vov._tc_modules()[0]
    for layer in list(mod.layers) + [vov.stem[3:5]]:
        conv, norm = layer[0], layer[1]
        l1, beta = tcconv.bound_consts(conv, norm)
        w, b = tcconv.folded(conv, norm)
        for seed, amp in ((1, 1.0), (2, 37.5)):
            x = synth.tensor((1, conv.in_channels, 9, 11), 700 + seed, -amp, amp)
            with torch.no_grad():
                y = F.relu(F.conv2d(x, w, b, conv.stride, conv.padding))
            assert float(y.max()) <= l1 * float(x.abs().max()) + beta
        # the worst case: every input of the heaviest filter's window at +-amax with the filter's signs
        with torch.no_grad():
            c = int(w.abs().sum((1, 2, 3)).argmax())
            x = torch.sign(w[c]).unsqueeze(0) * 3.0                      # [1, Cin, 3, 3]
            y = F.conv2d(x, w, b)                                        # one output pixel, no padding
            assert float(y.abs().max()) <= l1 * 3.0 + beta
            assert float(y[0, c]) - float(b[c] if b is not None else 0.0) >= 0.998 * (l1 / 1.001) * 3.0   # the bound is tight


def test_stem1_bound_holds_for_any_uint8_image():
    model, vov = _vov()
    mean, std = model._mean_std_host()
    bound = float(vov._stem1_bound(mean, std))
    w, b = vov._stem1_folded()
    imgs = [synth.ore_image(64, 96, 5).float(), torch.zeros(3, 64, 96), torch.full((3, 64, 96), 255.0),
            (torch.arange(3 * 64 * 96).reshape(3, 64, 96) % 2 * 255).float()]
    m, s = torch.tensor(mean).view(1, 3, 1, 1), torch.tensor(std).view(1, 3, 1, 1)
    with torch.no_grad():
        for im in imgs:
            y = F.relu(F.conv2d((im.unsqueeze(0) - m) / s, w, b, stride=2, padding=1))
            assert float(y.max()) <= bound
    assert vov._stem1_bound_rows(mean, std, 4).shape == (1, 4)


def test_split_handoff_is_offered_only_for_whole_16_channel_groups():
    _, vov = _vov()
    mods = vov._tc_modules()
    assert all(m._split_eligible() for m in mods)          # VoVNet-19-slim: slices of 64 / 80 / 96 / 112 channels
    assert vov.stem_u8_writes_split()
    mods[1].layers[0][0].out_channels = 72                 # not a multiple of 16 any more
    assert not mods[1]._split_eligible()
