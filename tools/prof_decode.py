"""Where the decode op's time goes: the same 64-problem, stride-8/16/32 pyramid with the selection and the emission
switched off in turn (pre_topk above the pixel count = no radix select; threshold near 1 = no candidates)."""
import torch
from faster_orefsdet_b200 import ops

dev = torch.device("cuda:0")
P, sizes, strides = 64, (80, 40, 20), (8, 16, 32)
g = torch.Generator(device="cpu").manual_seed(0)
taps = [(torch.randn((P, 48, s, s), generator=g) * 0.3).to(dev).contiguous(memory_format=torch.channels_last) for s in sizes]
status = torch.zeros(1, dtype=torch.int32, device=dev)


def run(label, thresh, pre_topk):
    cap = 3 * max(pre_topk, 1)
    for _ in range(3):
        ops.decode_topk_taps(taps, [0.0] * 5, strides, thresh, pre_topk, status, reg_scale=[1.0] * 3, cand_cap=cap)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()       # host cost of the wrapper out of the picture
    with torch.cuda.graph(graph):
        for _ in range(20):
            out = ops.decode_topk_taps(taps, [0.0] * 5, strides, thresh, pre_topk, status, reg_scale=[1.0] * 3, cand_cap=cap)
    graph.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    graph.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"{label:42s} {s.elapsed_time(e) / 20 * 1000:8.1f} us   count[0]={int(out[4][0])}")


run("default (thresh 1e-4, top 1000 per level)", 1e-4, 1000)
run("no select (top 6400: every pixel emitted)", 1e-4, 6400)
run("no candidates (thresh 0.999999)", 0.999999, 1000)
run("select, few emitted (top 32)", 1e-4, 32)
