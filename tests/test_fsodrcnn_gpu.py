"""GPU parity of the FsodRCNN path against outputs recorded from the UNMODIFIED reference
(tests/golden/make_golden_fsodrcnn.py -> fsodrcnn.npz): correlation gate, RPN logits, proposals, relation logits /
deltas, final detections after the class-wise NMS and the rescale, 1-way and 2-way."""
import numpy as np
import pytest
import torch

from faster_orefsdet_b200 import synth
from tests.util import golden, t


def fsodrcnn_support(class_ids, seed):
    """Same generator as tests/golden/make_golden_fsodrcnn.py::fsodrcnn_support."""
    d = {"res4_avg": {}, "res5_avg": {}}
    for j, c in enumerate(class_ids):
        d["res4_avg"][c] = synth.tensor((1, 1024, 14, 14), seed * 31 + 2 * j, 0.0, 1.2)
        d["res5_avg"][c] = synth.tensor((1, 2048, 7, 7), seed * 31 + 2 * j + 1, 0.0, 1.0)
    return d


def _match(ref_boxes, ref_scores, boxes, scores, box_tol=2e-2, score_rtol=1e-4):
    if ref_boxes.shape[0] == 0:
        return 1.0
    d = (ref_boxes[:, None, :] - boxes[None, :, :]).abs().amax(-1)
    j = d.argmin(1)
    ok = (d[torch.arange(len(j)), j] < box_tol) & ((ref_scores - scores[j]).abs() <= score_rtol * ref_scores.abs() + 1e-6)
    return float(ok.float().mean())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_fsodrcnn_head_matches_unmodified_reference(tag):
    from tests.test_fsodrcnn_cpu import build, param_shapes
    g = golden("fsodrcnn")
    model = build().cuda()
    model.load_state_dict(synth.state_dict(param_shapes()))
    class_ids = [int(c) for c in g[f"{tag}_class_ids"]]
    h, w, oh, ow, feat_seed, sup_seed = [int(v) for v in g[f"{tag}_size"]]
    model.set_prototypes(fsodrcnn_support(class_ids, sup_seed))
    res4 = synth.tensor((1, 1024, (h + 15) // 16, (w + 15) // 16), feat_seed, 0.0, 2.0).cuda()
    (ob, os_, ocls, oc), tr = model.head({"res4": res4}, [(h, w)], [(oh, ow)])
    for ci in range(len(class_ids)):
        corr = (res4 * tr["gate"][ci].reshape(1, -1, 1, 1)).reshape(-1)[::53].cpu()
        assert torch.allclose(corr, t(g[f"{tag}_corr{ci}"]), rtol=1e-4, atol=1e-5)
        p = tr["proposals"][ci][0]
        rb, rl = t(g[f"{tag}_prop_boxes{ci}"]), t(g[f"{tag}_prop_logits{ci}"])
        assert abs(len(p) - rb.shape[0]) <= 2
        assert _match(rb, rl, p.proposal_boxes.tensor.cpu(), p.objectness_logits.cpu(), box_tol=5e-2) >= 0.97
        logits, deltas = tr["raw"][ci]
        n = min(len(p), rb.shape[0])
        # rows correspond when the proposal lists do: compare the rows whose proposal boxes coincide
        same = ((p.proposal_boxes.tensor.cpu()[:n] - rb[:n]).abs().amax(1) < 5e-2)
        assert float(same.float().mean()) >= 0.95
        rl2, rd2 = t(g[f"{tag}_cls_logits{ci}"])[:n][same], t(g[f"{tag}_deltas{ci}"])[:n][same]
        assert torch.allclose(logits[:n].cpu()[same], rl2, rtol=1e-3, atol=2e-3)
        assert torch.allclose(deltas[:n].cpu()[same], rd2, rtol=1e-3, atol=2e-3)
    m = int(oc[0])
    rb, rs, rc = t(g[f"{tag}_out_boxes"]), t(g[f"{tag}_out_scores"]), t(g[f"{tag}_out_classes"])
    assert abs(m - rb.shape[0]) <= 2
    gb, gs, gc = ob[0, :m].cpu(), os_[0, :m].cpu(), ocls[0, :m].cpu()
    assert set(gc.tolist()) <= set(class_ids)
    matched = 0.0
    for c in class_ids:
        k = int((rc == c).sum())
        if k:
            matched += _match(rb[rc == c], rs[rc == c], gb[gc == c], gs[gc == c], box_tol=5e-2, score_rtol=1e-3) * k
    assert matched / max(rb.shape[0], 1) >= 0.95
    assert torch.all(gs[:-1] >= gs[1:])


@pytest.mark.gpu
def test_fsodrcnn_full_forward_batch_and_pkl(tmp_path, monkeypatch):
    """model(batched_inputs) through the real ResNet, support features from ./support_dir/support_feature.pkl; a batch of
    two images gives what two single-image calls give; pred_classes carry the support class ids as int8."""
    import os
    import pickle
    from tests.test_fsodrcnn_cpu import build, param_shapes
    monkeypatch.chdir(tmp_path)
    os.makedirs("support_dir")
    with open("support_dir/support_feature.pkl", "wb") as f:
        pickle.dump(fsodrcnn_support([3, 9], 11), f)
    model = build().cuda()
    model.load_state_dict(synth.state_dict(param_shapes()))
    imgs = [synth.ore_image(160, 192, 1000), synth.ore_image(160, 192, 1001)]
    out = model([{"image": imgs[0], "height": 200, "width": 240}, {"image": imgs[1]}])
    assert out[0]["instances"].image_size == (200, 240) and out[1]["instances"].image_size == (160, 192)
    assert out[0]["instances"].pred_classes.dtype == torch.int8
    assert set(out[0]["instances"].pred_classes.tolist()) <= {3, 9}
    for i, inp in enumerate(({"image": imgs[0], "height": 200, "width": 240}, {"image": imgs[1]})):
        single = model([inp])[0]["instances"]
        both = out[i]["instances"]
        assert abs(len(single) - len(both)) <= 2
        assert _match(single.pred_boxes.tensor.cpu(), single.scores.cpu(), both.pred_boxes.tensor.cpu(), both.scores.cpu(),
                      box_tol=5e-2, score_rtol=1e-3) >= 0.95
