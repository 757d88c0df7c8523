"""Phase cycle counts of nms_proposals_kernel (CTA 0), needs a -DFOD_NMS_PROF build: FOD_B200_LIB_DEV=build/libfod_nmsprof.so"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faster_orefsdet_b200 import ops, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2400
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
post = int(sys.argv[4]) if len(sys.argv) > 4 else 256
dev = "cuda"
torch.manual_seed(0)
status = ops.new_status(dev)
if os.environ.get("REAL", "0") == "1":      # the detector's own candidates on synthetic features (bench weights)
    from faster_orefsdet_b200 import synth
    from faster_orefsdet_b200.config import get_cfg
    from faster_orefsdet_b200.modeling import build_model
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs/fsod/finetune_vovnet.yaml"))
    cfg.merge_from_list(["MODEL.DEVICE", "cuda", "MODEL.CENTERNET.PRE_NMS_TOPK_TEST", n // 3 if n > 3000 else 1000, "MODEL.CENTERNET.POST_NMS_TOPK_TEST", post])
    model = build_model(cfg).eval()
    model.load_state_dict(synth.state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model.set_prototypes(synth.prototypes([1], 25, 7))
    H, W = (800, 1344) if n > 3000 else (640, 640)
    imgs = torch.stack([synth.ore_image(H, W, 1000 + i % 8) for i in range(P)]).cuda()
    with torch.no_grad():
        feats = model.features_from_uint8(imgs)
        _, tr = model.head(feats, [(H, W)] * P, [(H, W)] * P, want_trace=True)
    pr = tr["proposals"]
    boxes, scores, count = pr.cand_boxes.contiguous(), pr.cand_scores.contiguous(), pr.cand_count
    cap = boxes.shape[1]
    n = int(count[0])
    print("real candidates:", count[:4].tolist(), "proposals:", pr.count[:4].tolist())
else:
    ctr = torch.rand(P, cap, 2, device=dev) * 600 + 20
    wh = torch.rand(P, cap, 2, device=dev) * 60 + 100
    boxes = torch.cat((ctr - wh / 2, ctr + wh / 2), -1).contiguous()
    scores = torch.rand(P, cap, device=dev)
    count = torch.full((P,), n, dtype=torch.int32, device=dev)
L = _lib.lib()
roi_cap = (post + 64 + 63) // 64 * 64
for _ in range(3):
    ops.nms_proposals(boxes, scores, count, 0.6, post, roi_cap, status)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
has_prof = hasattr(L, 'fod_nms_prof')      # only in a -DFOD_NMS_PROF build
if has_prof:
    L.fod_nms_prof(buf, 1)
iters = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    keep, ob, os_, oc = ops.nms_proposals(boxes, scores, count, 0.6, post, roi_cap, status)
e1.record()
torch.cuda.synchronize()
if has_prof:
    L.fod_nms_prof(buf, 0)
v = [x / iters for x in buf]
print(f"P={P} n={n} post={post}: {e0.elapsed_time(e1)/iters*1e3:.1f} us/call; kept(sweep) {v[8]:.0f}, out {int(oc[0])}")
print(f"  chunks {v[9]:.0f}  fixed-point rounds {v[10]:.0f}  step2 cumulative: after fixed point {v[11]:.0f}, after kept writes {v[12]:.0f}, after lane-0 tail {v[13]:.0f}")
print(f"  cycles: load+sort+gather {v[6]:.0f}  sweep {v[7]:.0f}  [step1 {v[1]:.0f} step2 {v[2]:.0f} step3 {v[3]:.0f} push+sync {v[4]:.0f} loop-total {v[5]:.0f}]")
print(f"  step3 of thread 0 (cumulative): after colmask {v[14]:.0f}, after its suppression passes {v[15]:.0f}")
