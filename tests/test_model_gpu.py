"""GPU parity of the assembled head / detector against (a) golden outputs of the unmodified
reference and (b) the CPU oracle on the same features."""
import os
import pickle

import numpy as np
import pytest
import torch

from faster_orefsdet_b200 import ops, synth
from faster_orefsdet_b200.config import get_cfg
from faster_orefsdet_b200.modeling import build_model
from oracle import head_oracle as O
from tests.util import assert_close, golden, head_param_shapes, head_state_dict, t

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = O.HeadConfig()


def _model(*opts):
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs/fsod/finetune_vovnet.yaml"))
    cfg.merge_from_list(["MODEL.DEVICE", "cuda"] + list(opts))
    torch.manual_seed(0)
    model = build_model(cfg).eval()
    model.load_state_dict(synth.state_dict(head_param_shapes()), strict=False)
    return model


def _match(ref_boxes, ref_scores, boxes, scores, box_tol=2e-2, score_rtol=1e-4):
    """Fraction of reference detections that have a counterpart (same box within tol, score within rtol)."""
    if ref_boxes.shape[0] == 0:
        return 1.0
    d = (ref_boxes[:, None, :] - boxes[None, :, :]).abs().amax(-1)
    j = d.argmin(1)
    ok = (d[torch.arange(len(j)), j] < box_tol) & ((ref_scores - scores[j]).abs() <= score_rtol * ref_scores.abs() + 1e-6)
    return float(ok.float().mean())


@pytest.mark.parametrize("name", ["full_small", "full_640"])
def test_head_matches_unmodified_reference_outputs(name):
    g = golden(name)
    model = _model()
    model.set_prototypes(synth.prototypes(list(g["class_ids"]), int(g["shots"]), int(g["proto_seed"])))
    for i, ((h, w), (oh, ow)) in enumerate(zip(g["sizes"], g["out_sizes"])):
        feats = {k: v.cuda() for k, v in synth.features(1, int(h), int(w), int(g["feat_seed"]) + i).items()}
        (ob, os_, ocls, oc), tr = model.head(feats, [(int(h), int(w))], [(int(oh), int(ow))], want_trace=True)
        for l in range(3):
            ga = g[f"img{i}_attn{l}"]
            a = tr["attn"][l][0].cpu()
            assert_close(a if ga.ndim == 3 else a.reshape(-1)[::97], t(ga), what=f"attn{l}")
        n = int(tr["proposals"].count[0])
        rp, ro = t(g[f"img{i}_proposal_boxes"]), t(g[f"img{i}_objectness"])
        assert abs(n - rp.shape[0]) <= 2
        assert _match(rp, ro, tr["proposals"].boxes[0, :n].cpu(), tr["proposals"].scores[0, :n].cpu()) >= 0.98
        m = int(oc[0])
        rb, rs = t(g[f"img{i}_out_boxes"]), t(g[f"img{i}_out_scores"])
        assert abs(m - rb.shape[0]) <= 2
        assert _match(rb, rs, ob[0, :m].cpu(), os_[0, :m].cpu()) >= 0.97
        assert torch.all(os_[0, :m - 1] >= os_[0, 1:m])
        assert torch.all(ocls[0, :m] == 0)


def test_head_stagewise_bit_exact_indices_against_oracle():
    """Teacher-forced: each index-producing stage gets the oracle's inputs, so keep / top-k indices
    must be identical (north_star: bit-exact given identical scores)."""
    sd = head_state_dict()
    model = _model()
    protos = synth.prototypes([4], 5, 3)
    model.set_prototypes(protos)
    feats = synth.features(1, 320, 384, 77)
    tr = {}
    O.detect_image(feats, protos, sd, (320, 384), CFG, None, tr)
    pc = tr["per_class"][0]
    status = ops.new_status("cuda")
    hm = [x.cuda() for x in pc["hm"]]
    reg = [x.cuda() for x in pc["reg"]]
    prob = [x.sigmoid().cuda() for x in pc["hm"]]
    boxes, scores, loc, lc, cc = ops.decode_topk(prob, reg, (8, 16, 32), CFG.inference_th, 1000, status, hm_is_logit=False)
    n = int(cc[0])
    assert np.array_equal(loc[0, :n].cpu().numpy(), torch.cat(pc["loc"]).numpy())
    assert np.array_equal(boxes[0, :n].cpu().numpy(), pc["cand_boxes"].numpy())
    cb = torch.zeros((1, 3000, 4)); cs = torch.zeros((1, 3000))
    cb[0, :n], cs[0, :n] = pc["cand_boxes"], pc["cand_scores"]
    keep, pb, ps, pcount = ops.nms_proposals(cb.cuda(), cs.cuda(), cc, CFG.nms_th, CFG.post_nms_topk, 320, status)
    m = int(pcount[0])
    assert np.array_equal(keep[0, :m].cpu().numpy(), pc["keep"].numpy())
    db = torch.zeros((1, 320, 4)); ds = torch.zeros((1, 320))
    r = pc["det_boxes"].shape[0]
    db[0, :r], ds[0, :r] = pc["det_boxes"], pc["det_scores"]
    hw = torch.tensor([[320, 384]], dtype=torch.int32, device="cuda")
    ob, os_, ocls, orow, oc = ops.final_detect(db.cuda(), ds.cuda(), torch.tensor([r], dtype=torch.int32, device="cuda"), 1,
                                               CFG.score_thresh_test, CFG.nms_thresh_test, 100, hw, None, status)
    ops.check_status(status)
    k = int(oc[0])
    assert np.array_equal(orow[0, :k].cpu().numpy(), tr["final_rows"].numpy())


def test_nway_batched_head_matches_oracle():
    sd = head_state_dict()
    model = _model()
    class_ids = [7, 2, 9]
    protos = synth.prototypes(class_ids, 4, 11)
    model.set_prototypes(protos)
    B, H, W = 3, 256, 320
    feats = synth.features(B, H, W, 91)
    (ob, os_, ocls, oc) = model.head({k: v.cuda() for k, v in feats.items()}, [(H, W)] * B, [(H, W), (300, 400), (128, 160)])
    outs = [(H, W), (300, 400), (128, 160)]
    for b in range(B):
        fb = {k: v[b:b + 1] for k, v in feats.items()}
        rb, rs, rc = O.detect_image(fb, protos, sd, (H, W), CFG, outs[b])
        m = int(oc[b])
        assert abs(m - rb.shape[0]) <= 2
        gb, gs, gc = ob[b, :m].cpu(), os_[b, :m].cpu(), ocls[b, :m].cpu()
        assert set(gc.tolist()) <= {0, 1, 2}
        matched = 0.0
        for c in range(3):       # match class by class: different classes may propose the same box
            if int((rc == c).sum()) == 0:
                continue
            assert int((gc == c).sum()) > 0
            matched += _match(rb[rc == c], rs[rc == c], gb[gc == c], gs[gc == c]) * int((rc == c).sum())
        assert matched / rb.shape[0] >= 0.97


def test_highres_topk2000_head_matches_oracle():
    """BASELINE.json configs[3]: 1333x800 queries (padded 800x1344, M = 22 050), PRE/POST_NMS_TOPK_TEST 2000, two
    classes: top-k is exercised on all three levels, <= 6000 candidates per problem enter the NMS, <= 2000 ROIs per
    problem the relation head, and the class-wise final NMS sees both classes."""
    sd = head_state_dict()
    model = _model("MODEL.CENTERNET.PRE_NMS_TOPK_TEST", 2000, "MODEL.CENTERNET.POST_NMS_TOPK_TEST", 2000)
    cfg = O.HeadConfig(pre_nms_topk=2000, post_nms_topk=2000)
    class_ids = [3, 8]
    protos = synth.prototypes(class_ids, 5, 17)
    model.set_prototypes(protos)
    H, W = 800, 1344
    feats = synth.features(1, H, W, 123)
    (ob, os_, ocls, oc), tr = model.head({k: v.cuda() for k, v in feats.items()}, [(H, W)], [(H, W)], want_trace=True)
    otr = {}
    rb, rs, rc = O.detect_image(feats, protos, sd, (H, W), cfg, None, otr)
    # proposals per class: same count (+- ties at the post-NMS threshold) and the same boxes / objectness
    for c in range(2):
        pc = otr["per_class"][c]
        n = int(tr["proposals"].count[c])
        assert pc["cand_boxes"].shape[0] > 4000            # every level contributed its top-k
        assert abs(n - pc["proposals"].shape[0]) <= 2
        assert _match(pc["proposals"], pc["objectness"], tr["proposals"].boxes[c, :n].cpu(), tr["proposals"].scores[c, :n].cpu()) >= 0.98
    m = int(oc[0])
    assert abs(m - rb.shape[0]) <= 2
    gb, gs, gc = ob[0, :m].cpu(), os_[0, :m].cpu(), ocls[0, :m].cpu()
    matched = 0.0
    for c in range(2):
        if int((rc == c).sum()) == 0:
            continue
        matched += _match(rb[rc == c], rs[rc == c], gb[gc == c], gs[gc == c]) * int((rc == c).sum())
    assert matched / max(rb.shape[0], 1) >= 0.97


def _match_by_class(rb, rs, rc, gb, gs, gc, num_classes):
    """Fraction of oracle detections reproduced, matched class by class (different classes may propose the same box)."""
    if rb.shape[0] == 0:
        return 1.0
    matched = 0.0
    for c in range(num_classes):
        k = int((rc == c).sum())
        if k == 0:
            continue
        if int((gc == c).sum()) == 0:
            continue
        matched += _match(rb[rc == c], rs[rc == c], gb[gc == c], gs[gc == c]) * k
    return matched / rb.shape[0]


def test_config3_ten_way_ten_shot_head_matches_oracle():
    """BASELINE.json configs[2]: 10 support classes x 10 shots against a batch of 640x640 queries (B = 4 here so that
    the CPU oracle stays within seconds; the batch-32 run is timed by bench.py).  10 classes x 3 levels = 30 tap sets,
    i.e. five launches of the correlation kernel with class_begin > 0, ten per-class relation biases, and up to
    10 x 320 rows per image in the class-wise final NMS.  Every (image, class) proposal list and the final detections
    of two images are compared with the oracle."""
    sd = head_state_dict()
    model = _model()
    class_ids = [11, 3, 7, 2, 19, 5, 13, 17, 23, 29]
    C = len(class_ids)
    protos = synth.prototypes(class_ids, 10, 19)
    model.set_prototypes(protos)
    B, H, W = 4, 640, 640
    feats = synth.features(B, H, W, 191)
    (ob, os_, ocls, oc), tr = model.head({k: v.cuda() for k, v in feats.items()}, [(H, W)] * B, [(H, W)] * B, want_trace=True)
    assert tr["proposals"].count.shape[0] == B * C
    for b in (0, 3):
        fb = {k: v[b:b + 1] for k, v in feats.items()}
        otr = {}
        rb, rs, rc = O.detect_image(fb, protos, sd, (H, W), CFG, None, otr)
        for c in range(C):
            pc = otr["per_class"][c]
            p = b * C + c
            n = int(tr["proposals"].count[p])
            assert abs(n - pc["proposals"].shape[0]) <= 2, (b, c, n, pc["proposals"].shape[0])
            assert _match(pc["proposals"], pc["objectness"], tr["proposals"].boxes[p, :n].cpu(),
                          tr["proposals"].scores[p, :n].cpu()) >= 0.98, (b, c)
            # per-ROI scores of this class: the class's own folded bias was used
            assert _match(pc["det_boxes"], pc["det_scores"], tr["det_boxes"][p, :n].cpu(), tr["det_scores"][p, :n].cpu()) >= 0.97, (b, c)
        m = int(oc[b])
        assert abs(m - rb.shape[0]) <= 2
        gb, gs, gc = ob[b, :m].cpu(), os_[b, :m].cpu(), ocls[b, :m].cpu()
        assert set(gc.tolist()) <= set(range(C))
        assert _match_by_class(rb, rs, rc, gb, gs, gc, C) >= 0.97


def test_config4_highres_batch16_topk2000_matches_oracle():
    """BASELINE.json configs[3] at its batch size: 16 queries of 1333x800 (padded 800x1344, M = 22 050),
    PRE/POST_NMS_TOPK_TEST 2000; images 0 and 15 against the oracle."""
    sd = head_state_dict()
    model = _model("MODEL.CENTERNET.PRE_NMS_TOPK_TEST", 2000, "MODEL.CENTERNET.POST_NMS_TOPK_TEST", 2000)
    cfg = O.HeadConfig(pre_nms_topk=2000, post_nms_topk=2000)
    protos = synth.prototypes([5], 5, 23)
    model.set_prototypes(protos)
    B, H, W = 16, 800, 1344
    feats = synth.features(B, H, W, 223)
    (ob, os_, ocls, oc), tr = model.head({k: v.cuda() for k, v in feats.items()}, [(H, W)] * B, [(H, W)] * B, want_trace=True)
    for b in (0, 15):
        otr = {}
        rb, rs, rc = O.detect_image({k: v[b:b + 1] for k, v in feats.items()}, protos, sd, (H, W), cfg, None, otr)
        pc = otr["per_class"][0]
        n = int(tr["proposals"].count[b])
        assert pc["cand_boxes"].shape[0] > 4000
        assert abs(n - pc["proposals"].shape[0]) <= 2
        assert _match(pc["proposals"], pc["objectness"], tr["proposals"].boxes[b, :n].cpu(), tr["proposals"].scores[b, :n].cpu()) >= 0.98
        m = int(oc[b])
        assert abs(m - rb.shape[0]) <= 2
        assert _match(rb, rs, ob[b, :m].cpu(), os_[b, :m].cpu()) >= 0.97


def test_batched_call_equals_separate_calls():
    """B >= 1 images per call must give, per image, exactly what B separate calls give.  Every operand scale
    of the head's tensor-core kernels is per problem / per image (correlation bounds, GroupNorm bounds, feature bounds of
    the relation head), so the results are bit-identical, not merely close."""
    model = _model()
    model.set_prototypes(synth.prototypes([1, 4], 5, 7))
    B, H, W = 4, 256, 320
    feats = {k: v.cuda() for k, v in synth.features(B, H, W, 55).items()}
    for k in feats:                       # very different magnitudes per image: a batch-wide scale would differ from a single one
        feats[k] = feats[k] * torch.tensor([1.0, 37.0, 0.02, 5.0], device="cuda").reshape(B, 1, 1, 1)
    ob, os_, ocls, oc = model.head(feats, [(H, W)] * B, [(H, W)] * B)
    for b in range(B):
        fb = {k: v[b:b + 1].contiguous(memory_format=torch.channels_last) for k, v in feats.items()}
        sb, ss, sc, scount = model.head(fb, [(H, W)], [(H, W)])
        assert int(scount[0]) == int(oc[b])
        m = int(oc[b])
        assert torch.equal(sb[0, :m], ob[b, :m]) and torch.equal(ss[0, :m], os_[b, :m]) and torch.equal(sc[0, :m], ocls[b, :m])


def test_full_size_batch64_properties():
    """BASELINE.json configs[1] at its full size (batch 64 x 640x640, 1-way 25-shot) through the whole detector, checked
    by size-independent properties: per-image results do not depend on the batch they ran in (sampled images re-run
    alone from the same features give bit-identical detections), detections are sorted by score, at most DETECTIONS_PER_IMAGE of them, inside the
    image, no surviving pair overlaps above NMS_THRESH_TEST, and the final NMS is idempotent on its own output."""
    from torchvision.ops import box_iou
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1], 25, 7))
    B, H, W = 64, 640, 640
    base = [synth.ore_image(H, W, 1000 + i) for i in range(8)]
    imgs = torch.stack([torch.roll(base[i % 8], shifts=(7 * (i // 8), 13 * (i // 8)), dims=(1, 2)) for i in range(B)]).cuda()
    with torch.no_grad():
        feats = model.features_from_uint8(imgs)
        ob, os_, ocls, oc = model.head(feats, [(H, W)] * B, [(H, W)] * B)
        status = ops.new_status("cuda")
        total = 0
        for b in range(B):
            m = int(oc[b])
            total += m
            assert 0 < m <= 100
            bx, sc = ob[b, :m], os_[b, :m]
            assert torch.all(sc[:-1] >= sc[1:]) and torch.all(ocls[b, :m] == 0)
            assert float(bx[:, 0::2].min()) >= 0 and float(bx[:, 0::2].max()) <= W and float(bx[:, 1::2].min()) >= 0 and float(bx[:, 1::2].max()) <= H
            iou = box_iou(bx, bx).triu(1)
            assert float(iou.max()) <= CFG.nms_thresh_test + 1e-4
            if b % 16 == 0:      # idempotence: the same NMS over its own output keeps every box, in the same order
                keep = ops.batched_nms(bx.contiguous(), sc.contiguous(), torch.zeros(m, dtype=torch.int64, device="cuda"),
                                       CFG.nms_thresh_test)
                assert keep.tolist() == list(range(m))
        assert total > B          # the synthetic scene produces detections everywhere
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        protos25 = synth.prototypes([1], 25, 7)
        for b in (0, 37, 63):     # the oracle on the features the model's own extractor produced for this image
            fb = {k: v[b:b + 1].cpu().contiguous() for k, v in feats.items()}
            rb, rs, rc = O.detect_image(fb, protos25, sd_cpu, (H, W), CFG)
            m = int(oc[b])
            assert abs(m - rb.shape[0]) <= 2, (b, m, rb.shape[0])
            assert _match(rb, rs, ob[b, :m].cpu(), os_[b, :m].cpu()) >= 0.97, b
        for b in (0, 37, 63):     # the same image alone, from the same features: bit-identical detections
            fb = {k: v[b:b + 1].contiguous(memory_format=torch.channels_last) for k, v in feats.items()}
            sb, ss, _, scount = model.head(fb, [(H, W)], [(H, W)])
            m, m1 = int(oc[b]), int(scount[0])
            assert m1 == m
            assert torch.equal(ob[b, :m], sb[0, :m1]) and torch.equal(os_[b, :m], ss[0, :m1])
    ops.check_status(status)


def test_whole_detector_batch_equals_single_image_calls():
    """model(batched_inputs) with B images == B calls with one image each, bit for bit, through the whole detector
    (feature extractor included): every power-of-two operand scale of the tensor-core kernels is per image."""
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1, 6], 5, 7))
    imgs = [synth.ore_image(192, 256, 4300 + i) for i in range(3)]
    imgs[1] = (imgs[1].float() * 0.25).to(torch.uint8)          # a much darker image: its own scales must be used
    batch = [{"image": im} for im in imgs]
    both = model(batch)
    for i in range(3):
        single = model([batch[i]])[0]["instances"]
        inst = both[i]["instances"]
        assert len(single) == len(inst) and len(inst) > 0
        assert torch.equal(single.pred_boxes.tensor, inst.pred_boxes.tensor)
        assert torch.equal(single.scores, inst.scores) and torch.equal(single.pred_classes, inst.pred_classes)


def test_full_detector_forward_with_pkl_side_channel(tmp_path, monkeypatch):
    """model(batched_inputs) through the real backbone, prototypes from ./support_dir/support_feature.pkl."""
    monkeypatch.chdir(tmp_path)
    os.makedirs("support_dir")
    protos = synth.prototypes([1], 5, 7)
    with open("support_dir/support_feature.pkl", "wb") as f:
        pickle.dump(protos, f)
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    imgs = [synth.ore_image(256, 320, 1000), synth.ore_image(224, 300, 1001)]
    inputs = [{"image": imgs[0], "height": 300, "width": 375}, {"image": imgs[1]}]
    out = model(inputs)
    assert len(out) == 2 and out[0]["instances"].image_size == (300, 375) and out[1]["instances"].image_size == (224, 300)
    # oracle on the features the model's own backbone produced
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    images = model.preprocess_image(inputs)
    with torch.no_grad():
        feats = {k: v.cpu().contiguous() for k, v in model.backbone(images.tensor).items()}
    for b, inp in enumerate(inputs):
        size = images.image_sizes[b]
        outsz = (inp.get("height", size[0]), inp.get("width", size[1]))
        rb, rs, rc = O.detect_image({k: v[b:b + 1] for k, v in feats.items()}, protos, sd, size, CFG, outsz)
        inst = out[b]["instances"]
        assert inst.pred_classes.dtype == torch.int64
        assert abs(len(inst) - rb.shape[0]) <= 2
        # Integration check through the real backbone.  The final score of a detection moves by a few 1e-4 relative
        # when its proposal box moves by 1e-3 px (ROIAlign -> K = 8192 relation head -> softmax), which is what fp32
        # rounding differences in the tower convolutions produce (cuDNN's fp32 engines and the 3xTF32 kernel are
        # equally far from float64, tools/dbg_tower.py).  The stage-by-stage 1e-4 parity is pinned by the golden-vector
        # tests above and in test_ops_gpu.py; here the detections must correspond one to one.
        assert _match(rb, rs, inst.pred_boxes.tensor.cpu(), inst.scores.cpu(), score_rtol=1e-3) >= 0.97
    # second call: pickle untouched -> cached bank is reused
    bank = model._bank
    model(inputs)
    assert model._bank is bank
    # the reduced episode was written next to the pickle as a binary file; a fresh process-like model loads that
    assert os.path.exists("support_dir/support_feature.fodb")
    model2 = _model()
    model2.load_state_dict(model.state_dict())
    import faster_orefsdet_b200.modeling.prototypes as P
    calls = []
    monkeypatch.setattr(P.SupportCache, "load", lambda self: calls.append(1) or (_ for _ in ()).throw(AssertionError("unpickled")))
    out2 = model2(inputs)
    assert not calls
    for a, b in zip(out, out2):
        assert torch.equal(a["instances"].scores, b["instances"].scores)
        assert torch.equal(a["instances"].pred_boxes.tensor, b["instances"].pred_boxes.tensor)


def test_missing_pkl_builds_cache_and_exits_like_the_reference(tmp_path, monkeypatch):
    import pandas as pd
    from PIL import Image
    monkeypatch.chdir(tmp_path)
    os.makedirs("datasets/coco/img")
    rows = []
    boxes = [[40.0, 30.0, 200.0, 180.0], [10.0, 60.0, 120.0, 250.0], [100.0, 100.0, 300.0, 240.0]]
    for s in range(3):
        img = synth.ore_image(256, 320, 500 + s).permute(1, 2, 0).numpy()
        Image.fromarray(img[:, :, ::-1].copy()).save(f"datasets/coco/img/s{s}.png")
        rows.append({"category_id": 1, "file_path": f"img/s{s}.png", "support_box": boxes[s]})
    pd.DataFrame(rows).to_pickle("datasets/coco/10_shot_support_df.pkl")
    model = _model("INPUT.FS.SUPPORT_SHOT", 3)
    with pytest.raises(SystemExit):
        model([{"image": synth.ore_image(64, 64, 1)}])
    with open("support_dir/support_feature.pkl", "rb") as f:
        d = pickle.load(f)
    assert set(d) == {"p3", "p4", "p5", "rcnn_8", "rcnn_4"}
    assert tuple(d["p3"][1].shape) == (1, 128, 32, 32) and tuple(d["rcnn_8"][1].shape) == (3, 128, 8, 8)
    assert tuple(d["rcnn_4"][1].shape) == (3, 128, 4, 4)


def test_prototype_builder_matches_reference_cache_build():
    """P1+P2 on the reference's own init_model output (tests/golden/prototypes.npz): same synthetic
    backbone maps in, same five pkl entries out."""
    g = golden("prototypes")
    model = _model()
    feats = synth.features(3, 256, 320, 900)

    class Stub(torch.nn.Module):
        size_divisibility = 32

        def forward(self, x):
            return {k: v.cuda().contiguous(memory_format=torch.channels_last) for k, v in feats.items()}

        def output_shape(self):
            return model_backbone.output_shape()

    model_backbone = model.backbone
    model.backbone = Stub()
    imgs = [synth.ore_image(256, 320, 500 + s) for s in range(3)]
    d = model.build_support_dict({1: imgs}, {1: [list(map(float, b)) for b in g["support_boxes"]]})
    for k in ("p3", "p4", "p5", "rcnn_8", "rcnn_4"):
        assert_close(d[k][1], t(g[k]), what=k)


def test_pipelined_uint8_input_path_equals_generic_path():
    """model(batched_inputs) with equally sized uint8 images (chunked host-to-device copies, normalisation fused into
    the stem kernel) must give the features of preprocess_image + backbone."""
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    imgs = [synth.ore_image(128, 160, 3000 + i) for i in range(5)]
    inputs = [{"image": im} for im in imgs]
    model.PIPELINE_CHUNK = 2
    feats, sizes = model._features_pipelined(inputs)
    assert feats is not None and sizes == [(128, 160)] * 5
    ref = model.backbone(model.preprocess_image(inputs).tensor)
    torch.cuda.synchronize()
    # stem_1 runs as fp32 FMAs on this path and as a tensor-core convolution over im2col rows on the other.  The
    # synthetic (hash) weights make the 20-layer extractor ill-conditioned: per-layer differences of 5e-7 of the output
    # scale grow to ~1e-4 of the map's maximum at p5 (tools/dbg_backbone.py; cuDNN's own fp32 engines end 3e-5 away from
    # float64 on the same network).
    for k in ref:
        assert_close(feats[k], ref[k], rtol=1e-3, atol=4e-4 * float(ref[k].abs().max()), what=k)
    # float images or a size that needs padding fall back to the generic path
    assert model._features_pipelined([{"image": imgs[0].float()}])[0] is None
    assert model._features_pipelined([{"image": synth.ore_image(100, 160, 1)}])[0] is None


def test_submitted_batches_equal_plain_calls():
    """model.submit(next) before model(pending current): a batch enqueued behind the one in flight must not disturb it,
    and a PendingBatch must give the detections of the plain call bit for bit (host mirror and device fields)."""
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1], 5, 7))
    model.PIPELINE_CHUNK = 2
    batches = [[{"image": synth.ore_image(128, 160, 4000 + 10 * b + i).pin_memory(), "height": 100 + b, "width": 150} for i in range(3)]
               for b in range(3)]
    plain = [model(b) for b in batches]
    nxt = model.submit(batches[0])
    staged = []
    for k in range(3):
        cur = nxt
        if k + 1 < 3:
            nxt = model.submit(batches[k + 1])
        assert type(cur).__name__ == "PendingBatch" and len(cur) == 3
        staged.append(model(cur))
    for rp, rs in zip(plain, staged):
        for a, b in zip(rp, rs):
            ia, ib = a["instances"], b["instances"]
            assert ia.image_size == ib.image_size
            assert torch.equal(ia.pred_boxes.tensor, ib.pred_boxes.tensor) and torch.equal(ia.scores, ib.scores)
            assert torch.equal(ia.pred_classes, ib.pred_classes)
            ha, hb = ia.to("cpu"), ib.to("cpu")
            assert torch.equal(ha.pred_boxes.tensor, hb.pred_boxes.tensor) and torch.equal(ha.scores, ib.scores.cpu())
    # inputs outside the uint8 fast path come back unchanged (nothing to prefetch) and still run
    odd = [{"image": synth.ore_image(100, 160, 1)}]
    assert model.submit(odd) is odd
    assert len(model(model.submit(odd))) == 1
    # two batches in flight through the ring of three input buffers; a third submit is refused until one is collected
    pend = [model.submit(batches[k % 3]) for k in range(2)]
    with pytest.raises(Exception, match="in flight"):
        model.submit(batches[2])
    outs = []
    for k in range(2, 6):
        outs.append(model(pend.pop(0)))
        pend.append(model.submit(batches[k % 3]))
    outs += [model(p) for p in pend]
    for k, r in enumerate(outs):
        for a, b in zip(plain[k % 3], r):
            assert torch.equal(a["instances"].scores, b["instances"].scores)


def test_cuda_graph_replay_equals_eager_launches():
    """detect_from_uint8: stem eager + one graph replay must equal the eager path bit for bit, also when the requested
    output sizes change between replays and when the batch content changes."""
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1], 5, 7))
    sizes = [(128, 160)] * 3
    for rnd, outs in enumerate(([(128, 160)] * 3, [(256, 320), (128, 160), (64, 80)])):
        x = torch.stack([synth.ore_image(128, 160, 4000 + 10 * rnd + i) for i in range(3)]).cuda()
        model.USE_CUDA_GRAPH = True
        got = [t.clone() for t in model.detect_from_uint8(x, sizes, outs)]
        model.USE_CUDA_GRAPH = False
        ref = model.detect_from_uint8(x, sizes, outs)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    assert model._graph["graph"] is not None
    # new weights invalidate the captured graph (it holds packed copies of the old ones)
    old_graph = model._graph["graph"]
    sd = {k: (v * 1.25 if k.endswith("fpn_output3.weight") else v) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    model.USE_CUDA_GRAPH = True
    got = [t.clone() for t in model.detect_from_uint8(x, sizes, outs)]
    assert model._graph["graph"] is not old_graph
    model.USE_CUDA_GRAPH = False
    for a, b in zip(got, model.detect_from_uint8(x, sizes, outs)):
        assert torch.equal(a, b)


def test_submodule_weight_reload_refreshes_bias_and_graph():
    """weights replaced through a SUB-module (model.roi_heads.load_state_dict) or edited in place must
    invalidate the captured graph and the folded per-class bias of a bank installed with set_prototypes."""
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1, 2], 5, 7))
    sizes = [(128, 160)] * 2
    x = torch.stack([synth.ore_image(128, 160, 4100 + i) for i in range(2)]).cuda()
    model.USE_CUDA_GRAPH = True
    first = [t.clone() for t in model.detect_from_uint8(x, sizes, sizes)]
    g0, bias0 = model._graph["graph"], model._bank.bias_cls.clone()
    sd = {k: v * 1.5 if k in ("conv2.weight", "conv2.bias", "box_head.0.fc1.bias") else v for k, v in model.roi_heads.state_dict().items()}
    model.roi_heads.load_state_dict(sd)                       # sub-module path: the model's own load_state_dict is not involved
    model.init_model()
    got = [t.clone() for t in model.detect_from_uint8(x, sizes, sizes)]
    assert model._graph["graph"] is not g0
    assert not torch.equal(model._bank.bias_cls, bias0)
    assert_close(model._bank.bias_cls, model.roi_heads.class_bias(model._bank.support_mean), rtol=0, atol=0, what="bias")
    model.USE_CUDA_GRAPH = False
    for a, b in zip(got, model.detect_from_uint8(x, sizes, sizes)):
        assert torch.equal(a, b)
    assert not torch.equal(got[1], first[1])                  # the scores moved with the weights
    # in-place edit of a tower weight: new graph as well
    model.USE_CUDA_GRAPH = True
    g1 = model._graph["graph"] if model._graph["key"] == model._graph_key(2, 128, 160) else None
    model.detect_from_uint8(x, sizes, sizes)
    g1 = model._graph["graph"]
    with torch.no_grad():
        model.proposal_generator.centernet_head.agn_hm.bias.add_(0.25)
    got = [t.clone() for t in model.detect_from_uint8(x, sizes, sizes)]
    assert model._graph["graph"] is not g1
    model.USE_CUDA_GRAPH = False
    for a, b in zip(got, model.detect_from_uint8(x, sizes, sizes)):
        assert torch.equal(a, b)


def test_abandoned_batch_frees_its_ring_slot():
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    model.set_prototypes(synth.prototypes([1], 5, 7))
    batch = [{"image": synth.ore_image(128, 160, 4200 + i).pin_memory()} for i in range(2)]
    ref = model(batch)
    a, b = model.submit(batch), model.submit(batch)
    with pytest.raises(Exception, match="in flight"):
        model.submit(batch)
    model.abandon(a)
    c = model.submit(batch)
    with pytest.raises(Exception, match="collected or abandoned"):
        model(a)
    for pend in (b, c):
        for x, y in zip(ref, model(pend)):
            assert torch.equal(x["instances"].scores, y["instances"].scores)
    # device-resident uint8 inputs go through the same staging path, ordered behind their producer on the current stream
    dev_batch = [{"image": (d["image"].cuda() + 0)} for d in batch]
    for x, y in zip(ref, model(dev_batch)):
        assert torch.equal(x["instances"].scores, y["instances"].scores)


def test_tensor_core_backbone_matches_reference_vovnet_fpn():
    """The feature extractor on the tensor-core path against the outputs recorded from the reference's own
    VoVNet-19-slim-eSE + FPN (tests/golden/backbone.npz; the CPU test_host_cpu.py checks the same vectors through ATen).
    Tolerance: 4e-4 of each map's maximum (see the note in the test above)."""
    from tests.util import golden, t
    model = _model()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("backbone.")}
    model.load_state_dict(synth.state_dict(shapes), strict=False)
    g = golden("backbone")
    x = synth.tensor((2, 3, 96, 160), 601, -120.0, 130.0).cuda().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        out = model.backbone(x)
    for k in ("p3", "p4", "p5"):
        ref = t(g[k])
        assert_close(out[k], ref, rtol=1e-3, atol=4e-4 * float(ref.abs().max()), what=k)
