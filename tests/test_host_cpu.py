"""CPU tests of the host side: config, registries, state_dict names, backbone parity with
the reference's own VoVNet+FPN, C-ABI symbol export, loud failure without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from faster_orefsdet_b200 import _lib, ops, synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import (META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY,
                                           PrototypeBank)
from tests.util import GOLDEN, assert_close, golden, head_param_shapes, t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(*opts):
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs/fsod/finetune_vovnet.yaml"))
    cfg.merge_from_list(list(opts))
    return cfg


def test_config_resolves_like_the_reference_log():
    cfg = _cfg()
    c = cfg.MODEL.CENTERNET
    assert cfg.MODEL.META_ARCHITECTURE == "CenterNet2Detector" and cfg.MODEL.PROPOSAL_GENERATOR.NAME == "CenterNet"
    assert (c.INFERENCE_TH, c.PRE_NMS_TOPK_TEST, c.POST_NMS_TOPK_TEST, c.NMS_TH_TEST) == (1e-5, 1000, 256, 0.6)
    assert c.ONLY_PROPOSAL and c.WITH_AGN_HM and list(c.FPN_STRIDES) == [8, 16, 32]
    assert cfg.MODEL.ROI_HEADS.NAME == "CustomCascadeROIHeads" and cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST == 0.9
    assert cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST == 0.0 and cfg.TEST.DETECTIONS_PER_IMAGE == 100
    assert cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION == 8 and cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION2 == 4
    assert tuple(cfg.MODEL.ROI_BOX_CASCADE_HEAD.IOUS) == (0.6,)
    assert cfg.INPUT.FS.SUPPORT_WAY == 1 and cfg.INPUT.FS.SUPPORT_SHOT == 24
    assert cfg.SOLVER.STEPS == (10000, 11000)
    cfg2 = _cfg("INPUT.FS.SUPPORT_SHOT", "25", "MODEL.CENTERNET.POST_NMS_TOPK_TEST", 2000)
    assert cfg2.INPUT.FS.SUPPORT_SHOT == 25 and cfg2.MODEL.CENTERNET.POST_NMS_TOPK_TEST == 2000
    with pytest.raises(KeyError):
        _cfg("MODEL.NO_SUCH_KEY", 1)


def test_registries_hold_the_reference_names():
    assert "CenterNet2Detector" in META_ARCH_REGISTRY and "FsodRCNN" in META_ARCH_REGISTRY
    assert "CenterNet" in PROPOSAL_GENERATOR_REGISTRY
    assert "CustomCascadeROIHeads" in ROI_HEADS_REGISTRY


def test_state_dict_names_and_shapes_match_the_reference():
    model = META_ARCH_REGISTRY.get("CenterNet2Detector")(_cfg())
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    ref = head_param_shapes()
    with open(os.path.join(GOLDEN, "backbone_param_shapes.txt")) as f:
        for line in f:
            p = line.split()
            ref[p[0]] = tuple(int(x) for x in p[1:])
    assert mine == ref


def test_cpu_device_is_refused():
    with pytest.raises(_lib.FodError):
        META_ARCH_REGISTRY.get("CenterNet2Detector")(_cfg("MODEL.DEVICE", "cpu"))


def test_backbone_matches_reference_vovnet_fpn():
    model = META_ARCH_REGISTRY.get("CenterNet2Detector")(_cfg()).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    g = golden("backbone")
    x = synth.tensor((2, 3, 96, 160), 601, -120.0, 130.0)
    with torch.no_grad():
        out = model.backbone(x)
    for k in ("p3", "p4", "p5"):
        assert_close(out[k], t(g[k]), rtol=1e-4, atol=1e-4, what=k)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include/fod_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t) (fod_\w+)\(", hdr, flags=re.M))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().fod_version() == 100


def test_ops_fail_loudly_on_cpu_tensors():
    with pytest.raises(_lib.FodError):
        ops.batched_nms(torch.zeros(4, 4), torch.zeros(4), None, 0.5)
    with pytest.raises(_lib.FodError):
        ops.support_taps(torch.zeros(1, 128, 8, 8))


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.lib()
    assert lib.fod_nms_proposals(None, None, None, 1, 10, 0.5, 1, 1, None, None, None, None, None, None) == -1
    assert "null pointer" in _lib.last_error()
    assert lib.fod_batched_nms(None, None, None, 100000, 0.5, None, ctypes.c_void_p(8), None) == -2


def test_prototype_bank_pack_roundtrip():
    C = 3
    bank = PrototypeBank(list(range(C)), [synth.tensor((C, 7, 128), 1 + l) for l in range(3)],
                         synth.tensor((C, 128, 8, 8), 9), synth.tensor((C, 128), 10))
    buf = bank.pack()
    assert buf.numel() == PrototypeBank.packed_numel(C)
    back = PrototypeBank.unpack(buf, bank.class_ids)
    for a, b in zip(bank.taps, back.taps):
        assert torch.equal(a, b)
    assert torch.equal(bank.support_mean, back.support_mean) and torch.equal(bank.bias_cls, back.bias_cls)


def test_binary_prototype_file_round_trip(tmp_path):
    """the reduced episode as a binary, mmap-able file (header + class ids + the NCCL broadcast payload)."""
    from faster_orefsdet_b200.modeling.prototypes import PrototypeBank, load_bank, save_bank
    C = 3
    bank = PrototypeBank([7, 2, 9], [synth.tensor((C, 7, 128), 10 + l, -1, 1) for l in range(3)],
                         synth.tensor((C, 128, 8, 8), 20, -1, 1), synth.tensor((C, 128), 21, -1, 1))
    path = str(tmp_path / "support_feature.fodb")
    save_bank(bank, path, (123456789, 4242))
    assert os.path.getsize(path) == 64 + 64 + 4 * PrototypeBank.packed_numel(C)
    got = load_bank(path, "cpu", (123456789, 4242))
    assert got.class_ids == [7, 2, 9]
    for a, b in zip(got.taps + [got.support_mean, got.bias_cls], bank.taps + [bank.support_mean, bank.bias_cls]):
        assert torch.equal(a, b)
    assert load_bank(path, "cpu", (123456789, 4243)) is None          # derived from another pickle
    assert load_bank(str(tmp_path / "missing.fodb"), "cpu") is None
    with open(path, "r+b") as f:                                       # truncated file is refused
        f.truncate(1000)
    assert load_bank(path, "cpu") is None


def test_detection_instances_host_mirror_is_used_only_while_the_fields_are_untouched():
    """Instances.to("cpu") may return the host copy made by the detector's single transfer only while the
    fields are the ones it was built with."""
    from faster_orefsdet_b200.compat import Boxes, Instances
    from faster_orefsdet_b200.modeling.roi_heads import DetectionInstances, mirrored_instances

    def make():
        dev = {"pred_boxes": Boxes(torch.arange(8.0).reshape(2, 4)), "scores": torch.tensor([0.9, 0.8]),
               "pred_classes": torch.zeros(2, dtype=torch.int64)}
        host = {"pred_boxes": Boxes(torch.arange(8.0).reshape(2, 4) + 100), "scores": torch.tensor([0.9, 0.8]) + 100,
                "pred_classes": torch.zeros(2, dtype=torch.int64)}
        return mirrored_instances((10, 12), dev, host)

    inst = make()
    assert isinstance(inst, Instances) and isinstance(inst, DetectionInstances) and len(inst) == 2
    for target in ("cpu", torch.device("cpu")):
        h = inst.to(target)
        assert float(h.scores[0]) == pytest.approx(100.9) and h.image_size == (10, 12) and type(h) is Instances
    assert float(inst.to(device="cpu", non_blocking=True).scores[0]) == pytest.approx(100.9)
    # any other target, a replaced / added / removed field or an in-place edit: the plain field-by-field path
    assert float(inst.to(torch.float64 if False else "cpu", ).scores[0]) == pytest.approx(100.9)
    inst.scores = torch.tensor([0.5, 0.4])
    assert float(inst.to("cpu").scores[0]) == pytest.approx(0.5)
    inst = make()
    inst.pred_masks = torch.zeros(2, 3)
    assert inst.to("cpu").has("pred_masks") and float(inst.to("cpu").scores[0]) == pytest.approx(0.9)
    inst = make()
    inst.pred_boxes.tensor.mul_(2.0)          # Boxes.scale() edits in place
    assert float(inst.to("cpu").pred_boxes.tensor[1, 3]) == pytest.approx(14.0)
    inst = make()
    inst.remove("pred_classes")
    assert not inst.to("cpu").has("pred_classes")
