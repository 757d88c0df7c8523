"""GPU parity of the tensor-core convolution (fod_conv2d_nhwc) against PyTorch's CPU convolution evaluated in
float64 (a floating-point kernel: the torch op is the reference; tolerance 2e-5 relative to the output scale,
well inside the 1e-4 of north_star; measured ~5e-7)."""
import pytest
import torch
import torch.nn.functional as F

from faster_orefsdet_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(x, w, b, relu, stride=1):
    y = F.conv2d(x.double(), w.double(), None if b is None else b.double(), stride=stride, padding=w.shape[-1] // 2)
    return (y.relu() if relu else y).float()


def _check(y, ref, what):
    scale = float(ref.abs().max()) + 1e-12
    err = float((y.cpu() - ref).abs().max())
    assert err <= 2e-5 * scale, f"{what}: max abs err {err:.3e} vs output scale {scale:.3e}"


@pytest.mark.parametrize("n,h,w,cin,cout,k,relu", [
    (2, 20, 20, 128, 128, 3, True),     # tower conv on p5 (partial tiles in both directions)
    (1, 40, 40, 128, 128, 3, False),
    (2, 16, 32, 64, 64, 3, True),       # stem-like
    (1, 24, 40, 128, 64, 3, True),
    (2, 20, 24, 112, 80, 3, True),      # stage 3: Cin not a multiple of 32, Cout not a multiple of 32
    (1, 20, 20, 256, 96, 3, True),
    (1, 10, 10, 384, 112, 3, True),
    (1, 40, 40, 320, 112, 1, True),     # OSA concat 1x1
    (1, 20, 20, 352, 256, 1, True),     # two output-channel groups
    (1, 10, 12, 720, 512, 1, True),     # four groups, K tail chunk
    (3, 8, 16, 128, 128, 1, False),     # exactly one tile per image, odd tile count
    (1, 3, 5, 64, 128, 3, True),        # map smaller than a tile
    (1, 20, 20, 128, 8, 3, False),      # narrow output (agn_hm + bbox_pred padded to 8)
])
def test_conv2d_nhwc_matches_fp32(n, h, w, cin, cout, k, relu):
    x = synth.tensor((n, cin, h, w), 100 + cin + cout, -1.0, 1.0)
    wt = synth.tensor((cout, cin, k, k), 200 + cin + cout, -1.0, 1.0) / (cin * k * k) ** 0.5
    b = synth.tensor((cout,), 300 + cout, -0.5, 0.5)
    ref = _ref(x, wt, b, relu)
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last)
    packed = ops.conv2d_pack(wt.to(DEV))
    y = ops.conv2d_nhwc(xg, packed, b.to(DEV), cout, k, relu)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    _check(y, ref, f"conv {cin}->{cout} k{k} {h}x{w}")


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 64, 64, 64, 128), (1, 40, 56, 64, 128), (2, 17, 23, 32, 64), (1, 16, 32, 128, 96)])
def test_conv2d_nhwc_stride2_matches_fp32(n, h, w, cin, cout):
    x = synth.tensor((n, cin, h, w), 500 + h, -1.0, 1.0)
    wt = synth.tensor((cout, cin, 3, 3), 501 + h, -1.0, 1.0) / (cin * 9) ** 0.5
    b = synth.tensor((cout,), 502, -0.5, 0.5)
    ref = _ref(x, wt, b, True, stride=2)
    y = ops.conv2d_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), ops.conv2d_pack(wt.to(DEV)), b.to(DEV), cout, 3,
                        True, stride=2)
    assert y.shape == ref.shape
    _check(y, ref, f"stride-2 conv {cin}->{cout} {h}x{w}")


def test_conv2d_nhwc_channel_slices_concat_in_place():
    """Input and output as channel slices of wider NHWC buffers (the OSA concat is never materialised by a copy)."""
    n, h, w = 2, 24, 20
    buf = torch.zeros((n, h, w, 64 + 80 + 80), device=DEV).permute(0, 3, 1, 2)
    x = synth.tensor((n, 64, h, w), 7, -1.0, 1.0)
    w1 = synth.tensor((80, 64, 3, 3), 8, -0.05, 0.05)
    w2 = synth.tensor((80, 80, 3, 3), 9, -0.05, 0.05)
    buf[:, :64] = x.to(DEV)
    bounds = ops.new_amax(DEV, 3)          # one max|.| scalar per slice: a producer reports it, the consumer scales by it
    bounds[0:1].copy_(ops.absmax(x.to(DEV)))
    ops.conv2d_nhwc(buf[:, :64], ops.conv2d_pack(w1.to(DEV)), None, 80, 3, True, out=buf[:, 64:144], x_amax=bounds[0:1],
                    y_amax=bounds[1:2])
    ops.conv2d_nhwc(buf[:, 64:144], ops.conv2d_pack(w2.to(DEV)), None, 80, 3, True, out=buf[:, 144:224], x_amax=bounds[1:2],
                    y_amax=bounds[2:3])
    torch.cuda.synchronize()
    r1 = _ref(x, w1, None, True)
    r2 = _ref(r1, w2, None, True)
    _check(buf[:, :64], x, "input slice untouched")
    _check(buf[:, 64:144], r1, "first slice")
    _check(buf[:, 144:224], r2, "second slice")
    assert abs(float(bounds[1]) - float(r1.abs().max())) <= 1e-5 * float(r1.abs().max())
    assert abs(float(bounds[2]) - float(r2.abs().max())) <= 1e-5 * float(r2.abs().max())


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3e4, 1e9])
def test_conv2d_nhwc_operand_scaling_is_magnitude_independent(scale):
    """The fp16 split works on x * 2^e: inputs far outside fp16's range, and loose (larger) bounds, give the same
    relative accuracy."""
    x = synth.tensor((1, 64, 16, 16), 71, -1.0, 1.0) * scale
    wt = synth.tensor((64, 64, 3, 3), 72, -1.0, 1.0) / (24.0 * scale ** 0.5)
    ref = _ref(x, wt, None, False)
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last)
    pk = ops.conv2d_pack(wt.to(DEV))
    _check(ops.conv2d_nhwc(xg, pk, None, 64, 3), ref, f"scale {scale}")
    loose = ops.absmax(xg) * 37.0
    _check(ops.conv2d_nhwc(xg, pk, None, 64, 3, x_amax=loose), ref, f"scale {scale}, loose bound")


def test_conv2d_nhwc_no_bias_accumulation_bias_free():
    """Long K (3x3, 384 channels = 108 chunks = 14 partial sums) with same-sign products: the tensor core's
    round-toward-zero accumulation would show as a one-sided error without the partial-sum scheme."""
    x = synth.tensor((1, 384, 16, 16), 21, 0.5, 1.0)
    wt = synth.tensor((112, 384, 3, 3), 22, 0.5, 1.0) / 3456.0
    ref = _ref(x, wt, None, False)
    y = ops.conv2d_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), ops.conv2d_pack(wt.to(DEV)), None, 112, 3)
    rel = ((y.cpu() - ref) / ref)[:, :, 2:-2, 2:-2]
    assert float(rel.abs().max()) < 5e-6, float(rel.abs().max())
    assert abs(float(rel.mean())) < 4e-6, float(rel.mean())


@pytest.mark.parametrize("n,h,w,relu", [(3, 20, 20, True), (2, 37, 53, False), (1, 80, 80, True)])
def test_group_norm_nhwc_matches_fp32(n, h, w, relu):
    x = synth.tensor((n, 128, h, w), 41 + h, -2.0, 3.0)
    gamma, beta = synth.tensor((128,), 42, 0.5, 1.5), synth.tensor((128,), 43, -0.5, 0.5)
    ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5)
    ref = (ref.relu() if relu else ref).float()
    y = ops.group_norm_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), 32, gamma.to(DEV), beta.to(DEV), 1e-5, relu)
    _check(y, ref, "group_norm")


def test_stem_patches_and_gated_maxpool_match_torch():
    x = synth.tensor((2, 3, 37, 50), 61, -2.0, 2.0)
    w = synth.tensor((64, 3, 3, 3), 62, -0.3, 0.3)
    ref = F.conv2d(x.double(), w.double(), stride=2, padding=1).float()
    p = ops.stem_patches(x.to(DEV).contiguous(memory_format=torch.channels_last))
    assert tuple(p.shape) == (2, 32, 19, 25)
    wk = torch.cat((w.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5)), 1).reshape(64, 32, 1, 1)
    y = ops.conv2d_nhwc(p, ops.conv2d_pack(wk.to(DEV)), None, 64, 1)
    _check(y, ref, "stem_1 through im2col rows")

    t = synth.tensor((2, 24, 21, 30), 63, -1.0, 1.0)
    gate = synth.tensor((2, 24), 64, 0.0, 1.0)
    ref = F.max_pool2d(t, 3, 2, ceil_mode=True) * gate[:, :, None, None]
    buf = torch.zeros((2, 10, 15, 40), device=DEV).permute(0, 3, 1, 2)
    ops.maxpool3x3s2_nhwc(t.to(DEV).contiguous(memory_format=torch.channels_last), gate.to(DEV), out=buf[:, 8:32])
    assert torch.equal(buf[:, 8:32].cpu(), ref) and float(buf[:, :8].abs().max()) == 0 and float(buf[:, 32:].abs().max()) == 0


def test_stem1_u8_matches_normalise_then_conv():
    x = (synth.tensor((2, 3, 38, 52), 81, 0.0, 255.99)).to(torch.uint8)
    mean, std = [103.53, 116.28, 123.675], [1.0, 57.4, 58.4]
    w = synth.tensor((64, 3, 3, 3), 82, -0.3, 0.3)
    b = synth.tensor((64,), 83, -1.0, 1.0)
    xn = (x.double() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    ref = F.conv2d(xn, w.double(), b.double(), stride=2, padding=1).relu().float()
    bound = ops.new_amax(DEV)
    y = ops.stem1_u8(x.to(DEV), mean, std, w.to(DEV), b.to(DEV), y_amax=bound)
    assert tuple(y.shape) == (2, 64, 19, 26) and y.stride(1) == 1
    _check(y, ref, "stem_1")
    assert abs(float(bound) - float(ref.max())) <= 1e-5 * float(ref.max())


@pytest.mark.parametrize("std", [[1.0, 1.0, 1.0], [1.0, 57.4, 58.4]])
@pytest.mark.parametrize("hw", [(38, 52), (37, 51), (640, 640)])
def test_stem1_u8_tensor_cores_match_normalise_then_conv(std, hw):
    """The same layer through the tensor-core stem kernel (uint8 gather into tensor memory), odd sizes included (the last
    output row / column reads one pixel of padding), more tiles than SMs at 640x640, per-image output bounds."""
    h, w_ = hw
    n = 2 if h < 100 else 3
    x = (synth.tensor((n, 3, h, w_), 81, 0.0, 255.99)).to(torch.uint8)
    mean = [103.53, 116.28, 123.675]
    w = synth.tensor((64, 3, 3, 3), 82, -0.3, 0.3)
    b = synth.tensor((64,), 83, -1.0, 1.0)
    xn = (x.double() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    ref = F.conv2d(xn, w.double(), b.double(), stride=2, padding=1).relu().float()
    w32 = torch.cat((w.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5)), 1).reshape(64, 32, 1, 1).contiguous()
    pk = ops.conv2d_pack(w32.to(DEV))
    bound = ops.new_amax(DEV, n)
    y = ops.stem1_u8_tc(x.to(DEV), mean, std, pk, b.to(DEV), y_amax=bound)
    assert tuple(y.shape) == (n, 64, (h - 1) // 2 + 1, (w_ - 1) // 2 + 1) and y.stride(1) == 1
    _check(y, ref, "stem_1 (tensor cores)")
    for i in range(n):
        assert abs(float(bound.view(-1)[i]) - float(ref[i].max())) <= 1e-5 * float(ref[i].max())
    # and against the FP32 FMA kernel of the same layer
    y2 = ops.stem1_u8(x.to(DEV), mean, std, w.to(DEV), b.to(DEV))
    assert float((y - y2).abs().max()) <= 2e-6 * float(y2.abs().max())
    # a second launch on the same data gives the same bits (no state survives a launch)
    assert torch.equal(y, ops.stem1_u8_tc(x.to(DEV), mean, std, pk, b.to(DEV)))


@pytest.mark.parametrize("std,hw", [([1.0, 1.0, 1.0], (70, 90)), ([57.375, 57.12, 58.395], (129, 93)), ([57.375, 57.12, 58.395], (320, 320))])
def test_stem1_split_output_feeds_the_next_convolution_without_conversion(std, hw):
    """stem_1 writing the operand format of stem_2 (fp16 hi / lo of y * 2^e, the scale from a weights-only bound) and
    stem_2 reading it pre-split: against float64 of the two layers, against the fp32 hand-off, per-image max(y) kept."""
    h, w_ = hw
    n = 2
    x = (synth.tensor((n, 3, h, w_), 181, 0.0, 255.99)).to(torch.uint8)
    mean = [103.53, 116.28, 123.675]
    w1 = synth.tensor((64, 3, 3, 3), 182, -0.3, 0.3)
    b1 = synth.tensor((64,), 183, -1.0, 1.0)
    w2 = synth.tensor((64, 64, 3, 3), 184, -0.08, 0.08)
    b2 = synth.tensor((64,), 185, -0.5, 0.5)
    xn = (x.double() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    r1 = F.conv2d(xn, w1.double(), b1.double(), stride=2, padding=1).relu()
    ref = F.conv2d(r1, w2.double(), b2.double(), padding=1).relu().float()
    w32 = torch.cat((w1.permute(0, 2, 3, 1).reshape(64, 27), torch.zeros(64, 5)), 1).reshape(64, 32, 1, 1).contiguous()
    pk1, pk2 = ops.conv2d_pack(w32.to(DEV)), ops.conv2d_pack(w2.to(DEV))
    xmax = torch.tensor([max(abs(0.0 - m), abs(255.0 - m)) / sd for m, sd in zip(mean, std)], dtype=torch.float64)
    bound = (((w1.double().abs().sum((2, 3)) * xmax.view(1, 3)).sum(1) + b1.double().abs()).max() * 1.001).float().reshape(1).to(DEV)
    assert float(bound) >= float(r1.max())
    a1 = ops.new_amax(DEV, n)
    y1 = ops.stem1_u8_tc(x.to(DEV), mean, std, pk1, b1.to(DEV), y_amax=a1, y_bound=bound)
    for i in range(n):
        assert abs(float(a1.view(-1)[i]) - float(r1[i].max())) <= 1e-5 * float(r1[i].max())
    y2 = ops.conv2d_nhwc(y1, pk2, b2.to(DEV), 64, 3, True, x_amax=bound, x_presplit=True)
    _check(y2, ref, "stem_2 on a pre-split stem_1")
    # the fp32 hand-off of the same two layers
    a1f = ops.new_amax(DEV, n)
    y1f = ops.stem1_u8_tc(x.to(DEV), mean, std, pk1, b1.to(DEV), y_amax=a1f)
    y2f = ops.conv2d_nhwc(y1f, pk2, b2.to(DEV), 64, 3, True, x_amax=a1f.view(1, n))
    assert float((y2 - y2f).abs().max()) <= 2e-6 * float(y2f.abs().max())
    # every image alone gives the same bits (the scale does not depend on the data)
    y1s = ops.stem1_u8_tc(x[1:].to(DEV), mean, std, pk1, b1.to(DEV), y_bound=bound)
    assert torch.equal(ops.conv2d_nhwc(y1s, pk2, b2.to(DEV), 64, 3, True, x_amax=bound, x_presplit=True)[0], y2[1])


@pytest.mark.parametrize("cin,sc,h,w", [(128, 64, 40, 56), (64, 96, 20, 28), (112, 80, 20, 28), (48, 112, 12, 20)])
def test_split_handoff_through_an_osa_block(cin, sc, h, w):
    """fod_conv2d_nhwc_split on an OSA-shaped block: three 3x3 layers write their slices of one concat buffer in the
    split operand format (the first reads fp32), each reads its predecessor pre-split, the concat 1x1 reads the fp32
    slice and the three pre-split slices (each at its own published scale); against float64 and against the fp32
    hand-off of the same block; images whose magnitudes differ 1 000-fold keep their own scales."""
    n, cc = 3, 112
    x = synth.tensor((n, cin, h, w), 201, -1.0, 1.0) * torch.tensor([1.0, 1e-3, 30.0]).view(n, 1, 1, 1)
    ws = [synth.tensor((sc, cin if i == 0 else sc, 3, 3), 202 + i, -0.05, 0.05) for i in range(3)]
    bs = [synth.tensor((sc,), 206 + i, -0.2, 0.2) for i in range(3)]
    wc = synth.tensor((cc, cin + 3 * sc, 1, 1), 210, -0.05, 0.05)
    bc = synth.tensor((cc,), 211, -0.3, 0.3)
    outs = [x.double()]
    for i in range(3):
        outs.append(F.conv2d(outs[-1], ws[i].double(), bs[i].double(), padding=1).relu())
    ref = F.conv2d(torch.cat(outs, 1), wc.double(), bc.double()).relu().float()

    def run(split, x=x, n=n):
        full = n == 3
        buf = torch.empty((n, h, w, cin + 3 * sc), dtype=torch.float32, device=DEV).permute(0, 3, 1, 2)
        buf[:, :cin].copy_(x.to(DEV))
        amax = torch.zeros((4, n), device=DEV)
        amax[0] = x.abs().amax((1, 2, 3)).to(DEV)
        act = torch.zeros((3, n), device=DEV)
        src, off = buf[:, :cin], cin
        for i in range(3):
            dst = buf[:, off:off + sc]
            pk = ops.conv2d_pack(ws[i].to(DEV))
            if split:
                l1 = float(ws[i].double().abs().sum((1, 2, 3)).max()) * 1.001
                beta = float(bs[i].abs().max()) * 1.001
                ops.conv2d_nhwc_split(src, pk, bs[i].to(DEV), sc, 3, dst, amax[i:i + 1], y_amax=act[i], x_presplit=i > 0,
                                      x_actual=act[i - 1] if i > 0 else None, y_bound=amax[i + 1], y_l1=l1, y_beta=beta)
            else:
                ops.conv2d_nhwc(src, pk, bs[i].to(DEV), sc, 3, True, out=dst, x_amax=amax[i:i + 1], y_amax=amax[i + 1])
            src, off = dst, off + sc
        y = torch.empty((n, h, w, cc), dtype=torch.float32, device=DEV).permute(0, 3, 1, 2)
        pkc = ops.conv2d_pack(wc.to(DEV))
        if split:
            ops.conv2d_nhwc_split(buf, pkc, bc.to(DEV), cc, 1, y, amax, x_presplit_from=cin, slice_ch=[0, cin, cin + sc, cin + 2 * sc])
            for i in range(3 if full else 0):      # the published bounds hold, the actual maxima are the true ones
                assert bool((amax[i + 1].cpu() >= outs[i + 1].amax((1, 2, 3)).float()).all())
                assert torch.allclose(act[i].cpu(), outs[i + 1].amax((1, 2, 3)).float(), rtol=1e-5)
        else:
            ops.conv2d_nhwc(buf, pkc, bc.to(DEV), cc, 1, True, out=y, x_amax=amax)
        return y

    y_split, y_fp32 = run(True), run(False)
    for i in range(n):          # per image: the magnitudes differ by orders
        _check(y_split[i:i + 1], ref[i:i + 1], f"split hand-off, image {i}")
        assert float((y_split[i] - y_fp32[i]).abs().max()) <= 4e-6 * float(y_fp32[i].abs().max())
    # an image alone gives the same bits as inside the batch
    assert torch.equal(run(True, x[2:3], 1)[0], y_split[2])


@pytest.mark.parametrize("c,h,w", [(112, 41, 56), (256, 20, 27)])
def test_maxpool_split_output_feeds_a_3x3_convolution(c, h, w):
    """The stage pooling writing the split hand-off format (bound = the maximum of its source, gate <= 1) and the 3x3
    layer that reads it pre-split, against float64 and against the fp32 hand-off; odd sizes (ceil mode)."""
    n, co = 2, 80
    x = synth.tensor((n, c, h, w), 221, 0.0, 3.0) * torch.tensor([1.0, 40.0]).view(n, 1, 1, 1)
    gate = synth.tensor((n, c), 222, 0.0, 1.0)
    wt = synth.tensor((co, c, 3, 3), 223, -0.05, 0.05)
    b = synth.tensor((co,), 224, -0.3, 0.3)
    pooled = F.max_pool2d(x.double(), 3, 2, ceil_mode=True) * gate.double().view(n, c, 1, 1)
    ref = F.conv2d(pooled, wt.double(), b.double(), padding=1).relu().float()
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    bound = x.amax((1, 2, 3)).to(DEV).contiguous()
    pk = ops.conv2d_pack(wt.to(DEV))
    ps = ops.maxpool3x3s2_nhwc(xd, gate.to(DEV), y_bound=bound)
    y = torch.empty((n, ref.shape[2], ref.shape[3], co), dtype=torch.float32, device=DEV).permute(0, 3, 1, 2)
    ops.conv2d_nhwc_split(ps, pk, b.to(DEV), co, 3, y, bound.view(1, n), x_presplit=True)
    _check(y, ref, "3x3 on a pre-split pooled map")
    pf = ops.maxpool3x3s2_nhwc(xd, gate.to(DEV))
    assert float((pf.cpu() - pooled.float()).abs().max()) <= 1e-6 * float(pooled.max())
    yf = ops.conv2d_nhwc(pf, pk, b.to(DEV), co, 3, True, x_amax=bound.view(1, n))
    for i in range(n):
        assert float((y[i] - yf[i]).abs().max()) <= 4e-6 * float(yf[i].abs().max())


@pytest.mark.parametrize("up2", [False, True])
def test_conv2d_nhwc_residual_in_epilogue(up2):
    """FPN top-down step: lateral 1x1 + (nearest 2x upsampled) coarser map, fused into the convolution's epilogue."""
    n, h, w, cin, cout = 2, 20, 28, 96, 128
    x = synth.tensor((n, cin, h, w), 91, -1.0, 1.0)
    wt = synth.tensor((cout, cin, 1, 1), 92, -0.1, 0.1)
    b = synth.tensor((cout,), 93, -0.5, 0.5)
    r = synth.tensor((n, cout, h // 2, w // 2) if up2 else (n, cout, h, w), 94, -2.0, 2.0)
    ref = _ref(x, wt, b, False) + (F.interpolate(r, scale_factor=2.0, mode="nearest") if up2 else r)
    bound = ops.new_amax(DEV)
    y = ops.conv2d_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), ops.conv2d_pack(wt.to(DEV)), b.to(DEV), cout, 1,
                        residual=r.to(DEV).contiguous(memory_format=torch.channels_last), residual_upsample2=up2, y_amax=bound)
    _check(y, ref, "conv + residual")
    assert abs(float(bound) - float(ref.abs().max())) <= 1e-5 * float(ref.abs().max())


def test_conv2d_nhwc_ese_gate_pieces():
    """eSE without a pass over the map: per-tile channel sums out of the epilogue -> gate kernel; gate multiplied into the
    consumer's input (vovnet.py eSEModule)."""
    n, h, w, cin, c = 3, 20, 28, 64, 96
    x = synth.tensor((n, cin, h, w), 95, -1.0, 1.0)
    wt = synth.tensor((c, cin, 3, 3), 96, -0.1, 0.1)
    fw, fb = synth.tensor((c, c, 1, 1), 97, -0.2, 0.2), synth.tensor((c,), 98, -1.0, 1.0)
    y_ref = _ref(x, wt, None, True)
    g_ref = (F.relu6(F.conv2d(y_ref.double().mean((2, 3), keepdim=True), fw.double(), fb.double()) + 3.0) / 6.0).float()
    colsum = torch.zeros((n, ops.conv2d_tiles_per_image(h, w), c), device=DEV)
    y = ops.conv2d_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), ops.conv2d_pack(wt.to(DEV)), None, c, 3, True,
                        colsum=colsum)
    gate = ops.ese_gate(colsum, h * w, fw.to(DEV), fb.to(DEV))
    _check(y, y_ref, "conv")
    _check(gate.view(n, c, 1, 1), g_ref, "gate")
    # consumer: 1x1 convolution of the gated map, gate applied to the A operand
    w2 = synth.tensor((128, c, 1, 1), 99, -0.1, 0.1)
    ref2 = _ref(y_ref * g_ref, w2, None, False)
    out = ops.conv2d_nhwc(y, ops.conv2d_pack(w2.to(DEV)), None, 128, 1, a_gate=gate)
    _check(out, ref2, "gated consumer")


def test_conv_groupnorm_relu_conv_without_materialising_the_normalised_map():
    """CenterNetHead tower: conv -> GroupNorm(32) -> ReLU -> conv; statistics from the first convolution's epilogue,
    normalisation applied to the second convolution's input operand (zero padding must stay zero)."""
    n, h, w, c = 3, 20, 28, 128
    x = synth.tensor((n, c, h, w), 111, -1.0, 1.0)
    w1 = synth.tensor((c, c, 3, 3), 112, -0.05, 0.05)
    b1 = synth.tensor((c,), 113, -0.5, 0.5)
    gamma, beta = synth.tensor((c,), 114, 0.5, 1.5), synth.tensor((c,), 115, -0.5, 0.5)
    w2 = synth.tensor((8, c, 3, 3), 116, -0.05, 0.05)
    b2 = synth.tensor((8,), 117, -1.0, 1.0)
    t_ref = F.conv2d(x.double(), w1.double(), b1.double(), padding=1)
    g_ref = F.group_norm(t_ref, 32, gamma.double(), beta.double(), 1e-5).relu()
    y_ref = F.conv2d(g_ref, w2.double(), b2.double(), padding=1).float()
    tiles = ops.conv2d_tiles_per_image(h, w)
    cs, cq = torch.zeros((n, tiles, c), device=DEV), torch.zeros((n, tiles, c), device=DEV)
    a_t = ops.new_amax(DEV)
    t = ops.conv2d_nhwc(x.to(DEV).contiguous(memory_format=torch.channels_last), ops.conv2d_pack(w1.to(DEV)), b1.to(DEV), c, 3,
                        y_amax=a_t, colsum=cs, colsumsq=cq)
    scale, shift, a_g = ops.group_norm_affine(cs, cq, h * w, 32, gamma.to(DEV), beta.to(DEV), 1e-5, x_amax=a_t)
    y = ops.conv2d_nhwc(t, ops.conv2d_pack(w2.to(DEV)), b2.to(DEV), 8, 3, x_amax=a_g, a_gate=scale, a_shift=shift, a_relu=True)
    _check(y, y_ref, "conv-GN-ReLU-conv")
    assert float(a_g) >= float(g_ref.max()) * (1 - 1e-5)
