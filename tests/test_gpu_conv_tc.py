"""tcgen05 / TMA implicit-GEMM convolution vs a float64 evaluation on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

import synth

pytestmark = pytest.mark.gpu

CASES = [
    # kind, cin, cout, k, stride, pad, outpad, H, W
    ("conv", 64, 128, 3, 1, 1, 0, 16, 32), ("conv", 128, 256, 3, 1, 1, 0, 13, 21), ("conv", 256, 128, 3, 1, 1, 0, 8, 16),
    ("conv", 32, 64, 1, 1, 0, 0, 24, 32), ("conv", 64, 128, 1, 1, 0, 0, 9, 17), ("conv", 32, 64, 3, 2, 1, 0, 26, 38),
    ("conv", 64, 64, 3, 2, 1, 0, 16, 32), ("convT", 128, 64, 3, 2, 1, 1, 7, 11), ("convT", 64, 32, 2, 2, 0, 0, 12, 16),
    ("convT", 128, 64, 2, 2, 0, 0, 8, 16),
]


def cl(t, dtype=torch.bfloat16):
    return t.to("cuda:0").to(dtype).contiguous(memory_format=torch.channels_last)


def close(a, b, atol, rtol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs() - rtol * b.abs()
    assert err.max().item() <= atol, f"{what}: max abs err {(a - b).abs().max().item():.3e}"


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_tc_conv_forward_and_backward_data(case):
    from spaa_b200 import ops
    from spaa_b200._lib import lib
    ops.invalidate_packed_weights()
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 3
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    x = synth.randn(1, "tc.x", (B, cin, H, W)).to(torch.bfloat16)
    w = synth.randn(2, "tc.w", spec.weight_shape(), (2.0 / (cin * k * k)) ** 0.5)
    b = synth.randn(3, "tc.b", (cout,), 0.1)
    wq = w.to(torch.bfloat16).double()              # the kernel rounds weights to bf16
    xd = x.double().requires_grad_(True)
    pre = F.conv2d(xd, wq, b.double(), stride, pad) if kind == "conv" else F.conv_transpose2d(xd, wq, b.double(), stride, pad, outpad)
    add = synth.randn(4, "tc.add", pre.shape, 0.5).to(torch.bfloat16)
    ref = F.relu(pre + add.double())
    n0 = ops.launch_count()
    got = ops.conv_forward(spec, cl(x), w.to("cuda:0"), b.to("cuda:0"), add=cl(add), epi=ops.EPI_RELU)
    assert got.dtype == torch.bfloat16 and got.is_contiguous(memory_format=torch.channels_last)
    close(got.float(), ref, 2e-2, 1e-2, "forward")                    # bf16 output rounding: 2^-8 relative
    # tight check in fp32: difference to the CUDA-core kernel on identical operands is accumulation order only
    ops.TC_ENABLED = False
    try:
        simt = ops.conv_forward(spec, cl(x), wq.float().to("cuda:0"), b.to("cuda:0"), add=cl(add), epi=ops.EPI_RELU)
    finally:
        ops.TC_ENABLED = True
    d = (got.float() - simt.float()).abs()
    assert (d > 0.05 * simt.float().abs() + 2e-2).sum().item() == 0, f"TC vs CUDA-core: {d.max().item():.3e}"
    # backward-data with skip sum, ReLU mask and the dual-mask second output
    cot = synth.randn(5, "tc.cot", pre.shape).to(torch.bfloat16)
    gx, = torch.autograd.grad((pre * cot.double()).sum(), xd)
    extra = synth.randn(6, "tc.extra", x.shape, 0.5).to(torch.bfloat16)
    m = synth.randn(7, "tc.m", x.shape).to(torch.bfloat16)
    m2 = synth.randn(8, "tc.m2", x.shape).to(torch.bfloat16)
    out2 = torch.empty_like(cl(x))
    dx = ops.conv_backward_data(spec, cl(cot), w.to("cuda:0"), (H, W), add=cl(extra), mask=cl(m), mask_mode=ops.MASK_POS, mask2=cl(m2), out2=out2)
    refb = (gx + extra.double()) * (m.double() > 0)
    scale = max(1.0, refb.abs().max().item())
    close(dx.float(), refb, 2e-2 * scale, 1e-2, "backward data")
    close(out2.float(), refb * (m2.double() > 0), 2e-2 * scale, 1e-2, "backward data out2")
    assert lib().spaa_conv_tc_supported is not None and ops.launch_count() > n0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_tc_padded_input_and_planar_output(dtype):
    """The boundary layers on the tensor-core path: 3/6-channel images zero-padded to 16 NHWC channels (conv1, conv1_s,
    backward of conv6) and fp32 NCHW outputs with <= 32 channels (conv6 forward, backward of conv1)."""
    from spaa_b200 import ops
    B, H, W = 2, 20, 36
    dev = "cuda:0"
    # conv1_s-like: 6 input channels living at channels 3..8 of a padded 16-channel tensor, stride 2
    spec = ops.ConvSpec("conv", 6, 32, 3, 2, 1)
    x16 = torch.zeros(B, 16, H, W)
    x16[:, 3:9] = synth.randn(11, "pd.x", (B, 6, H, W))
    x16 = x16.to(dtype)
    w = synth.randn(12, "pd.w", spec.weight_shape(), 0.2)
    b = synth.randn(13, "pd.b", (32,), 0.1)
    ref = F.relu(F.conv2d(x16[:, 3:9].double(), w.to(dtype).double(), b.double(), 2, 1))
    got = ops.conv_forward(spec, cl(x16, dtype), w.to(dev), b.to(dev), epi=ops.EPI_RELU, cin_offset=3)
    assert got.dtype == dtype
    close(got.float(), ref, 2e-2, 1e-2, "padded-input forward")
    # conv6-like: 32 -> 3 channels, fp32 NCHW output, residual + ReLU + clamp
    spec6 = ops.ConvSpec("conv", 32, 3, 3, 1, 1)
    x7 = synth.randn(14, "pd.x7", (B, 32, H, W)).to(dtype)
    w6 = synth.randn(15, "pd.w6", spec6.weight_shape(), 0.1)
    b6 = synth.randn(16, "pd.b6", (3,), 0.1)
    res = synth.randn(17, "pd.res", (1, 3, H, W), 0.3)
    ref6 = torch.clamp(F.relu(F.conv2d(x7.double(), w6.to(dtype).double(), b6.double(), 1, 1) + res.double()), max=1)
    probe = ops.set_probe(lambda kind, sp: kind.endswith("_tc"))
    got6 = ops.conv_forward(spec6, cl(x7, dtype), w6.to(dev), b6.to(dev), add=res.to(dev), epi=ops.EPI_RELU | ops.EPI_CLAMP_MAX1, out_dtype=torch.float32)
    assert got6.dtype == torch.float32 and got6.is_contiguous() and len(probe["events"]) == 1
    close(got6, ref6, 2e-3, 1e-3, "planar forward")
    # backward of conv6: padded 16-channel cotangent (3 real) -> 32-channel gradient with ReLU mask of an fp16/bf16 activation
    cot = torch.zeros(B, 16, H, W)
    cot[:, :3] = synth.randn(18, "pd.cot", (B, 3, H, W))
    cotq = cot.to(torch.bfloat16)
    xd = x7.double().requires_grad_(True)
    pre = F.conv2d(xd, w6.to(torch.bfloat16).double(), None, 1, 1)
    gx, = torch.autograd.grad((pre * cotq[:, :3].double()).sum(), xd)
    d7 = ops.conv_backward_data(spec6, cl(cotq), w6.to(dev), (H, W), mask=cl(x7, dtype), mask_mode=ops.MASK_POS, out_dtype=torch.bfloat16)
    close(d7.float(), gx * (x7.double() > 0), 3e-2, 1e-2, "padded backward of conv6")
    # backward of conv1: 32-channel bf16 gradient -> 3-channel fp32 NCHW (up = 2 parity phases)
    spec1 = ops.ConvSpec("conv", 3, 32, 3, 2, 1)
    w1 = synth.randn(19, "pd.w1", spec1.weight_shape(), 0.2)
    d1 = synth.randn(20, "pd.d1", (B, 32, H // 2, W // 2)).to(torch.bfloat16)
    xin = torch.zeros(B, 3, H, W, dtype=torch.double, requires_grad=True)
    pre1 = F.conv2d(xin, w1.to(torch.bfloat16).double(), None, 2, 1)
    gx1, = torch.autograd.grad((pre1 * d1.double()).sum(), xin)
    dx = ops.conv_backward_data(spec1, cl(d1), w1.to(dev), (H, W), out_dtype=torch.float32)
    ops.set_probe(None)
    assert dx.dtype == torch.float32 and len(probe["events"]) == 3
    close(dx, gx1, 2e-3 * max(1.0, gx1.abs().max().item()), 1e-3, "planar backward of conv1")
    # packed warp output and packed cotangent
    img = synth.rand(21, "pd.img", (B, 3, 24, 24))
    grid = (synth.rand(22, "pd.grid", (2, H, W)) * 2.2 - 1.1).to(dev)
    mask = (synth.rand(23, "pd.mask", (H * W,)) > 0.2).float().to(dev)
    s = synth.rand(24, "pd.s", (1, 3, H, W)).to(dev)
    xw = ops.grid_sample(img.to(dev), grid, clamp01=True, mask=mask)
    pk = ops.grid_sample_packed(img.to(dev), grid, dtype, clamp01=True, mask=mask, rough=s)
    want = torch.zeros(B, 16, H, W, device=dev)
    want[:, :3], want[:, 3:6], want[:, 6:9] = xw, s, xw * s
    close(pk.float(), want.to(dtype).float(), 1e-2, 0, "packed warp")
    g0, g1 = synth.randn(25, "pd.g0", (B, 3, H, W)).to(dev), synth.randn(26, "pd.g1", (B, 3, H, W)).to(dev)
    sel = torch.tensor([1, 0], dtype=torch.uint8, device=dev)
    act = (synth.rand(27, "pd.act", (B, 3, H, W)) * 1.4 - 0.2).clamp(0, 1).to(dev)
    outp = torch.empty((B, 16, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    ops.select_cotangent_packed(g0, g1, sel, act, ops.MASK_OPEN01, outp)
    wantc = torch.zeros(B, 16, H, W, device=dev)
    wantc[:, :3] = ops.select_cotangent(g0, g1, sel, act, ops.MASK_OPEN01, torch.empty_like(g0))
    close(outp.float(), wantc.to(torch.bfloat16).float(), 0, 0, "packed cotangent")


WG_CASES = [
    # kind, cin, cout, k, stride, pad, outpad, H, W
    ("conv", 128, 256, 3, 1, 1, 0, 20, 24), ("conv", 256, 128, 3, 1, 1, 0, 17, 9), ("conv", 64, 128, 3, 1, 1, 0, 16, 32),
    ("conv", 32, 64, 1, 1, 0, 0, 24, 32), ("conv", 32, 64, 3, 2, 1, 0, 26, 38), ("conv", 64, 64, 3, 2, 1, 0, 16, 32),
    ("convT", 128, 64, 3, 2, 1, 1, 7, 11), ("convT", 64, 32, 2, 2, 0, 0, 12, 16), ("conv", 32, 16, 3, 1, 1, 0, 20, 36),
    ("conv", 16, 32, 3, 2, 1, 0, 20, 36),
]


@pytest.mark.parametrize("mix", ["bf16", "fp16"])
@pytest.mark.parametrize("case", WG_CASES, ids=lambda c: "-".join(map(str, c)))
def test_tc_backward_weight(case, mix):
    """tcgen05 MN-major backward-weight kernel vs autograd in float64 on the same 16-bit-rounded operands."""
    from spaa_b200 import ops
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 3
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    adt = torch.float16 if mix == "fp16" else torch.bfloat16
    x = synth.randn(31, "wg.x", (B, cin, H, W)).to(adt)
    w = torch.zeros(spec.weight_shape(), dtype=torch.double, requires_grad=True)
    pre = F.conv2d(x.double(), w, None, stride, pad) if kind == "conv" else F.conv_transpose2d(x.double(), w, None, stride, pad, outpad)
    dy = synth.randn(32, "wg.dy", pre.shape).to(adt)
    gw, = torch.autograd.grad((pre * dy.double()).sum(), w)
    dw = torch.zeros(spec.weight_shape(), device="cuda:0")
    db = torch.zeros(cout, device="cuda:0")
    probe = ops.set_probe(lambda kd, sp: kd == "bwd_weight_tc")
    ops.conv_backward_weight(spec, cl(x, adt), cl(dy, adt), dw, db)
    ops.set_probe(None)
    assert len(probe["events"]) == 1, "the tensor-core backward-weight kernel was not used"
    scale = gw.abs().max().item()
    close(dw, gw, 2e-3 * scale, 1e-3, f"dW {case}")
    close(db, dy.double().sum((0, 2, 3)), 1e-2 * max(1.0, dy.double().sum((0, 2, 3)).abs().max().item()), 1e-3, "dbias")
    # accumulation semantics: a second call adds
    ops.conv_backward_weight(spec, cl(x, adt), cl(dy, adt), dw, None)
    close(dw, 2 * gw, 4e-3 * scale, 1e-3, "dW accumulates")


def test_tc_backward_weight_padded_channels():
    """conv1_s-like (6 real input channels at 3..8 of a 16-channel tensor) and conv6-like (3 real output-gradient channels of 16)."""
    from spaa_b200 import ops
    B, H, W = 2, 20, 36
    spec = ops.ConvSpec("conv", 6, 32, 3, 2, 1)
    x16 = torch.zeros(B, 16, H, W)
    x16[:, 3:9] = synth.randn(41, "wgp.x", (B, 6, H, W))
    x16 = x16.to(torch.bfloat16)
    w = torch.zeros(spec.weight_shape(), dtype=torch.double, requires_grad=True)
    pre = F.conv2d(x16[:, 3:9].double(), w, None, 2, 1)
    dy = synth.randn(42, "wgp.dy", pre.shape).to(torch.bfloat16)
    gw, = torch.autograd.grad((pre * dy.double()).sum(), w)
    dw = torch.zeros(spec.weight_shape(), device="cuda:0")
    ops.conv_backward_weight(spec, cl(x16), cl(dy), dw, None, x_offset=3)
    close(dw, gw, 2e-3 * gw.abs().max().item(), 1e-3, "padded-input dW")
    spec6 = ops.ConvSpec("conv", 32, 3, 3, 1, 1)
    x7 = synth.randn(43, "wgp.x7", (B, 32, H, W)).to(torch.bfloat16)
    w6 = torch.zeros(spec6.weight_shape(), dtype=torch.double, requires_grad=True)
    pre6 = F.conv2d(x7.double(), w6, None, 1, 1)
    cot = torch.zeros(B, 16, H, W)
    cot[:, :3] = synth.randn(44, "wgp.cot", (B, 3, H, W))
    cot = cot.to(torch.bfloat16)
    gw6, = torch.autograd.grad((pre6 * cot[:, :3].double()).sum(), w6)
    dw6 = torch.zeros(spec6.weight_shape(), device="cuda:0")
    db6 = torch.zeros(3, device="cuda:0")
    ops.conv_backward_weight(spec6, cl(x7), cl(cot), dw6, db6)
    close(dw6, gw6, 2e-3 * gw6.abs().max().item(), 1e-3, "padded-gradient dW")
    close(db6, cot[:, :3].double().sum((0, 2, 3)), 1e-2, 1e-3, "padded-gradient dbias")


# ---------------------------------------------------------------------------------------------------------
# bf16x3 split-precision ("fp32-accurate") mode: three bf16 parts per value, six part products per multiply, fp32 accumulation in TMEM
# ---------------------------------------------------------------------------------------------------------

def split_cl(t):
    """fp32 [B,C,H,W] -> the split-precision operand: bf16 [B,3C,H,W] channels-last, channels [h(C) | m(C) | l(C)]."""
    from spaa_b200 import ops
    h, m, l = ops.split3(t.to("cuda:0").float())
    return torch.cat((h, m, l), 1).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def unsplit(t):
    c = t.shape[1] // 3
    return t[:, :c].double() + t[:, c:2 * c].double() + t[:, 2 * c:].double()


def test_split3_parts_reconstruct_fp32():
    from spaa_b200 import ops
    v = synth.randn(1, "sp.v", (4, 8, 6, 5)) * torch.tensor([1e-6, 1e-3, 1.0, 37.0, 1e3, 1e-2, 5.0, 0.3]).view(1, 8, 1, 1)
    s = split_cl(v)
    assert (unsplit(s).float().cpu() - v).abs().max().item() == 0.0       # 3 x 8 significand bits hold an fp32 value exactly
    packed = ops.pack_nhwc16(v[:, :3].to("cuda:0"), v[:, 3:8].to("cuda:0"), torch.bfloat16, split=True)
    assert packed.shape == (4, 48, 6, 5)
    got = packed[:, 0:16].double() + packed[:, 16:32].double() + packed[:, 32:48].double()
    assert (got[:, :8].float().cpu() - v).abs().max().item() == 0.0 and got[:, 8:].abs().max().item() == 0.0


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_tc_split_precision_conv_forward_and_backward_data(case):
    """Against float64 on the SAME fp32 operands (no bf16 rounding of anything): the tensor-core result must be fp32-accurate."""
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 3
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    x = synth.randn(1, "tc.x", (B, cin, H, W))
    w = synth.randn(2, "tc.w", spec.weight_shape(), (2.0 / (cin * k * k)) ** 0.5)
    b = synth.randn(3, "tc.b", (cout,), 0.1)
    xd = x.double().requires_grad_(True)
    pre = F.conv2d(xd, w.double(), b.double(), stride, pad) if kind == "conv" else F.conv_transpose2d(xd, w.double(), b.double(), stride, pad, outpad)
    # Shapes beyond the networks' own: a stride-2 layer with >= 64 input channels loads four 20 KB parity planes per stage; with the split mode's
    # three-part residual ring on top the shared-memory plan does not fit.  The networks' stride-2 layers have <= 32 input channels on the side that
    # carries a residual (conv2 / conv2_s) and only a ReLU mask on the wide side (the backward of transConv1): those cases run "lean" here.
    lean_f = kind == "conv" and stride == 2 and cin >= 64
    lean_b = kind == "convT" and cin >= 128
    add = synth.randn(4, "tc.add", pre.shape, 0.5) * (0.0 if lean_f else 1.0)
    ref = F.relu(pre + add.double())
    probe = ops.set_probe(lambda kind_, spec_: kind_.endswith("_tc"))
    got = ops.conv_forward(spec, split_cl(x), w.to("cuda:0"), b.to("cuda:0"), add=None if lean_f else split_cl(add), epi=ops.EPI_RELU, split=True)
    assert len(probe["events"]) == 1, "the split-precision layer did not run on the tensor-core kernel"
    ops.set_probe(None)
    assert got.dtype == torch.bfloat16 and got.shape[1] == 3 * cout and got.is_contiguous(memory_format=torch.channels_last)
    err = (unsplit(got).cpu() - ref).abs().max().item()
    # the exact-fp32 CUDA-core kernel on the same operands: the yardstick for "fp32-accurate"
    simt = ops.conv_forward(spec, x.to("cuda:0"), w.to("cuda:0"), b.to("cuda:0"), add=None if lean_f else add.to("cuda:0"), epi=ops.EPI_RELU)
    err_simt = (simt.double().cpu() - ref).abs().max().item()
    print(f"split-precision fwd {case}: max abs err vs float64 {err:.2e} (CUDA-core fp32 kernel: {err_simt:.2e}, output scale {ref.abs().max().item():.2f})")
    assert err <= max(3 * err_simt, 2e-6 * ref.abs().max().item()), (err, err_simt)
    # backward-data with skip sum, ReLU mask and the dual-mask second output
    cot = synth.randn(5, "tc.cot", pre.shape)
    gx, = torch.autograd.grad((pre * cot.double()).sum(), xd)
    extra = synth.randn(6, "tc.extra", x.shape, 0.5)
    m, m2 = synth.randn(7, "tc.m", x.shape), synth.randn(8, "tc.m2", x.shape)
    out2 = torch.empty_like(split_cl(x))
    if lean_b:
        dx = ops.conv_backward_data(spec, split_cl(cot), w.to("cuda:0"), (H, W), mask=split_cl(m), mask_mode=ops.MASK_POS, split=True)
        extra, out2, m2 = torch.zeros_like(extra), dx, torch.ones_like(m2)
    else:
        dx = ops.conv_backward_data(spec, split_cl(cot), w.to("cuda:0"), (H, W), add=split_cl(extra), mask=split_cl(m), mask_mode=ops.MASK_POS,
                                    mask2=split_cl(m2), out2=out2, split=True)
    refb = (gx + extra.double()) * (m.double() > 0)
    scale = max(1.0, refb.abs().max().item())
    eb = (unsplit(dx).cpu() - refb).abs().max().item()
    eb2 = (unsplit(out2).cpu() - refb * (m2.double() > 0)).abs().max().item()
    print(f"split-precision bwd {case}: max abs err {eb:.2e} / out2 {eb2:.2e} (scale {scale:.2f})")
    assert eb <= 5e-6 * scale and eb2 <= 5e-6 * scale, (eb, eb2, scale)      # sums of up to 2 304 products of O(1) x O(0.03) factors in fp32


def test_tc_split_precision_padded_input_and_planar_output():
    """The boundary layers: conv1 / conv1_s read the 48-channel [x | s | x*s | 0] operand, conv6 writes fp32 planes (+ fp32 residual, ReLU, clamp),
    conv6's backward reads the padded 3-channel cotangent, conv1's backward writes fp32 planes."""
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    B, H, W = 2, 20, 24
    x3 = synth.randn(1, "sp.x3", (B, 3, H, W))
    s6 = synth.randn(2, "sp.s6", (B, 6, H, W))
    packed = ops.pack_nhwc16(x3.to("cuda:0"), s6.to("cuda:0"), torch.bfloat16, split=True)
    for name, spec, cin_off, src in (("conv1", ops.ConvSpec("conv", 3, 32, 3, 2, 1), 0, x3), ("conv1_s", ops.ConvSpec("conv", 6, 32, 3, 2, 1), 3, s6)):
        w = synth.randn(3, "sp.w" + name, spec.weight_shape(), 0.2)
        b = synth.randn(4, "sp.b" + name, (32,), 0.1)
        ref = F.relu(F.conv2d(src.double(), w.double(), b.double(), 2, 1))
        got = ops.conv_forward(spec, packed, w.to("cuda:0"), b.to("cuda:0"), epi=ops.EPI_RELU, cin_offset=cin_off, split=True)
        e = (unsplit(got).cpu() - ref).abs().max().item()
        assert e <= 2e-6 * max(1.0, ref.abs().max().item()), (name, e)
        # backward of the strided layer to fp32 planes
        cot = synth.randn(5, "sp.cot" + name, ref.shape)
        xd = src.double().requires_grad_(True)
        gx, = torch.autograd.grad((F.conv2d(xd, w.double(), None, 2, 1) * cot.double()).sum(), xd)
        dx = ops.conv_backward_data(spec, split_cl(cot), w.to("cuda:0"), (H, W), out_dtype=torch.float32, split=True)
        assert dx.dtype == torch.float32 and dx.shape == src.shape
        e = (dx.double().cpu() - gx).abs().max().item()
        assert e <= 3e-6 * max(1.0, gx.abs().max().item()), (name + " bwd", e)
    spec6 = ops.ConvSpec("conv", 32, 3, 3, 1, 1)
    x7 = synth.randn(6, "sp.x7", (B, 32, H, W)).relu()
    w6, b6 = synth.randn(7, "sp.w6", spec6.weight_shape(), 0.1), synth.randn(8, "sp.b6", (3,), 0.1)
    res1 = synth.randn(9, "sp.res1", (1, 3, H, W), 0.3)
    ref = torch.clamp(F.relu(F.conv2d(x7.double(), w6.double(), b6.double(), 1, 1) + res1.double()), max=1)
    got = ops.conv_forward(spec6, split_cl(x7), w6.to("cuda:0"), b6.to("cuda:0"), add=res1.to("cuda:0"), epi=ops.EPI_RELU | ops.EPI_CLAMP_MAX1,
                           out_dtype=torch.float32, split=True)
    assert got.dtype == torch.float32 and got.shape == (B, 3, H, W)
    assert (got.double().cpu() - ref).abs().max().item() <= 2e-6
    cot3 = synth.randn(10, "sp.cot3", (B, 3, H, W))
    d_pk = ops.pack_nhwc16(cot3.to("cuda:0"), None, torch.bfloat16, split=True)
    xd = x7.double().requires_grad_(True)
    gx, = torch.autograd.grad((F.conv2d(xd, w6.double(), None, 1, 1) * cot3.double()).sum(), xd)
    d7 = ops.conv_backward_data(spec6, d_pk, w6.to("cuda:0"), (H, W), mask=split_cl(x7), mask_mode=ops.MASK_POS, split=True)
    e = (unsplit(d7).cpu() - gx * (x7.double() > 0)).abs().max().item()
    assert e <= 3e-6 * max(1.0, gx.abs().max().item()), e


@pytest.mark.parametrize("case", WG_CASES, ids=lambda c: "-".join(map(str, c)))
def test_tc_split_precision_backward_weight(case):
    """Six launches of the tcgen05 backward-weight kernel on the bf16 parts vs float64 autograd on the same fp32 operands."""
    from spaa_b200 import ops
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 3
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    x = synth.randn(31, "wg.x", (B, cin, H, W))
    w = torch.zeros(spec.weight_shape(), dtype=torch.double, requires_grad=True)
    pre = F.conv2d(x.double(), w, None, stride, pad) if kind == "conv" else F.conv_transpose2d(x.double(), w, None, stride, pad, outpad)
    dy = synth.randn(32, "wg.dy", pre.shape)
    gw, = torch.autograd.grad((pre * dy.double()).sum(), w)
    dw = torch.zeros(spec.weight_shape(), device="cuda:0")
    db = torch.zeros(cout, device="cuda:0")
    ops.conv_backward_weight(spec, split_cl(x), split_cl(dy), dw, db, split=True)
    scale = gw.abs().max().item()
    e = (dw.double().cpu() - gw).abs().max().item()
    print(f"split-precision dW {case}: max abs err {e:.2e} of {scale:.2e}")
    assert e <= 3e-6 * scale, (e, scale)              # sums of up to 10^4 products: fp32-level
    ref_b = dy.double().sum((0, 2, 3))
    assert (db.double().cpu() - ref_b).abs().max().item() <= 1e-5 * max(1.0, ref_b.abs().max().item())


# ---------------------------------------------------------------------------------------------------------
# compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer.md): own out-of-bounds-write check with guard bands
# ---------------------------------------------------------------------------------------------------------

def _guarded(shape, dtype, channels_last=True, guard=4096):
    """A tensor of `shape` carved out of the middle of a larger buffer filled with a canary pattern; returns (view, check())."""
    n = 1
    for s_ in shape:
        n *= s_
    buf = torch.full((n + 2 * guard,), -8192.0, dtype=dtype, device="cuda:0")
    B, C, H, W = shape
    if channels_last:
        view = buf[guard:guard + n].view(B, H, W, C).permute(0, 3, 1, 2)
    else:
        view = buf[guard:guard + n].view(B, C, H, W)

    def check(what):
        torch.cuda.synchronize()
        lo, hi = buf[:guard].float(), buf[guard + n:].float()
        assert bool((lo == -8192.0).all()) and bool((hi == -8192.0).all()), f"{what}: wrote outside its output buffer"
    return view, check


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("case", [("conv", 64, 128, 3, 1, 1, 0, 13, 21), ("conv", 32, 64, 3, 2, 1, 0, 27, 37), ("convT", 64, 32, 2, 2, 0, 0, 11, 15),
                                  ("convT", 128, 64, 3, 2, 1, 1, 7, 11), ("conv", 32, 64, 1, 1, 0, 0, 9, 5)], ids=lambda c: "-".join(map(str, c)))
def test_tc_conv_writes_stay_inside_ragged_outputs(case, split):
    """Ragged sizes (tiles hang over every edge) with the outputs placed between canary guard bands: the persistent tcgen05 kernel's cooperative
    128-bit stores, the parity-phase stores of the up-sampling layers and the three-part stores of the split mode must not touch a byte outside."""
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B, np_ = 2, (3 if split else 1)
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    Ho, Wo = spec.out_hw(H, W)
    x = synth.randn(1, "gb.x", (B, cin, H, W))
    w = synth.randn(2, "gb.w", spec.weight_shape(), 0.05).to("cuda:0")
    xin = split_cl(x) if split else cl(x)
    out, chk = _guarded((B, cout * np_, Ho, Wo), torch.bfloat16)
    ops.conv_forward(spec, xin, w, None, out=out, epi=ops.EPI_RELU, split=split)
    chk("forward")
    assert torch.isfinite(out.float()).all() and float(out.float().abs().max()) < 1e3
    dy = synth.randn(3, "gb.dy", (B, cout, Ho, Wo))
    dyin = split_cl(dy) if split else cl(dy)
    dx, chk2 = _guarded((B, cin * np_, H, W), torch.bfloat16)
    m = split_cl(x) if split else cl(x)
    ops.conv_backward_data(spec, dyin, w, (H, W), out=dx, mask=m, mask_mode=ops.MASK_POS, split=split)
    chk2("backward-data")
    assert torch.isfinite(dx.float()).all()


def test_planar_and_pack_kernels_write_inside_their_outputs():
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    B, H, W = 2, 19, 23
    spec6 = ops.ConvSpec("conv", 32, 3, 3, 1, 1)
    x7 = synth.randn(6, "gb.x7", (B, 32, H, W)).relu()
    w6 = synth.randn(7, "gb.w6", spec6.weight_shape(), 0.1).to("cuda:0")
    for split in (False, True):
        out, chk = _guarded((B, 3, H, W), torch.float32, channels_last=False)
        ops.conv_forward(spec6, split_cl(x7) if split else cl(x7), w6, None, out=out, epi=ops.EPI_RELU | ops.EPI_CLAMP_MAX1, out_dtype=torch.float32, split=split)
        chk(f"planar conv6 forward (split={split})")
    pk, chk = _guarded((B, 48, H, W), torch.bfloat16)
    ops.pack_nhwc16(synth.randn(8, "gb.p", (B, 3, H, W)).to("cuda:0"), synth.randn(9, "gb.s", (B, 6, H, W)).to("cuda:0"), torch.bfloat16, split=True, out=pk)
    chk("pack_nhwc16 split")


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("case", [("conv", 128, 256, 3, 96, 104), ("conv", 256, 128, 3, 96, 104), ("conv", 64, 128, 3, 100, 100)], ids=lambda c: "-".join(map(str, c)))
def test_tc_conv_pair_mode_many_tiles(case, split):
    """Wide layers with streamed weights and >= 2 tiles per CTA run in PAIR mode (two tiles per weight pass, four / two accumulator buffers in TMEM;
    conv_tc.cu): 312-350 tiles on 148 persistent CTAs, so CTAs own 2 or 3 tiles (a pair plus a single), rings wrap and accumulator buffers are reused.
    Forward with residual + ReLU and backward-data with mask (+ second masked output) against the exact CUDA-core kernel on the same operands."""
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    kind, cin, cout, k, H, W = case
    B = 4
    spec = ops.ConvSpec(kind, cin, cout, k, 1, 1, 0)
    x = synth.randn(1, "pm.x", (B, cin, H, W))
    w = synth.randn(2, "pm.w", spec.weight_shape(), (2.0 / (cin * k * k)) ** 0.5).to("cuda:0")
    b = synth.randn(3, "pm.b", (cout,), 0.1).to("cuda:0")
    add = synth.randn(4, "pm.add", (B, cout, H, W), 0.5)
    if split:
        got = unsplit(ops.conv_forward(spec, split_cl(x), w, b, add=split_cl(add), epi=ops.EPI_RELU, split=True)).float()
        ref = ops.conv_forward(spec, x.to("cuda:0"), w, b, add=add.to("cuda:0"), epi=ops.EPI_RELU).float()
        tol_a, tol_r = 3e-5, 0.0
    else:
        xq, aq = x.to(torch.bfloat16), add.to(torch.bfloat16)
        got = ops.conv_forward(spec, cl(xq), w, b, add=cl(aq), epi=ops.EPI_RELU).float()
        ref = ops.conv_forward(spec, xq.float().to("cuda:0"), w.to(torch.bfloat16).float(), b, add=aq.float().to("cuda:0"), epi=ops.EPI_RELU).float()
        tol_a, tol_r = 3e-2, 1e-2
    err = (got - ref).abs() - tol_r * ref.abs()
    assert err.max().item() <= tol_a, f"forward: max abs err {(got - ref).abs().max().item():.3e}"
    # every image and every tile position must be right (a mis-paired tile would be wrong as a whole): per-image, per-16-row-band maxima
    band = (got - ref).abs().amax(dim=(1, 3))
    assert (band <= tol_a + tol_r * ref.abs().max()).all()
    dy = synth.randn(5, "pm.dy", (B, cout, H, W))
    m = synth.randn(6, "pm.m", (B, cin, H, W))
    if split:
        dx = unsplit(ops.conv_backward_data(spec, split_cl(dy), w, (H, W), mask=split_cl(m), mask_mode=ops.MASK_POS, split=True)).float()
        refb = ops.conv_backward_data(spec, dy.to("cuda:0"), w, (H, W), mask=m.to("cuda:0"), mask_mode=ops.MASK_POS).float()
        tol_a, tol_r = 1e-4, 0.0
    else:
        dq, mq = dy.to(torch.bfloat16), m.to(torch.bfloat16)
        dx = ops.conv_backward_data(spec, cl(dq), w, (H, W), mask=cl(mq), mask_mode=ops.MASK_POS).float()
        refb = ops.conv_backward_data(spec, dq.float().to("cuda:0"), w.to(torch.bfloat16).float(), (H, W), mask=mq.float().to("cuda:0"), mask_mode=ops.MASK_POS).float()
        tol_a, tol_r = 0.1, 1e-2
    errb = (dx - refb).abs() - tol_r * refb.abs()
    assert errb.max().item() <= tol_a, f"backward-data: max abs err {(dx - refb).abs().max().item():.3e}"


@pytest.mark.parametrize("case", [("conv", 32, 64, 3, 2, 1, 0, 26, 38), ("conv", 64, 128, 3, 1, 1, 0, 16, 32), ("convT", 64, 32, 2, 2, 0, 0, 12, 16),
                                  ("convT", 128, 64, 3, 2, 1, 1, 7, 11), ("conv", 32, 64, 1, 1, 0, 0, 24, 32)], ids=lambda c: "-".join(map(str, c)))
def test_tc_forward_bf16_copy_of_fp16_output(case):
    """SPAA_EPI_OUT2_BF16 (the fp16 training mode): the forward epilogue also writes the bf16 rounding of the SAME fp32 result -- both the
    register-lean epilogue of the narrow layers (TMA stores) and the epilogue of the wide ones; the fp16 output itself is unchanged."""
    from spaa_b200 import ops
    ops.invalidate_packed_weights()
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 3
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    x = cl(synth.randn(1, "c2.x", (B, cin, H, W)), torch.float16)
    w = synth.randn(2, "c2.w", spec.weight_shape(), (2.0 / (cin * k * k)) ** 0.5).to("cuda:0")
    b = synth.randn(3, "c2.b", (cout,), 0.1).to("cuda:0")
    Ho, Wo = spec.out_hw(H, W)
    add = cl(synth.randn(4, "c2.add", (B, cout, Ho, Wo), 0.5), torch.float16)
    plain = ops.conv_forward(spec, x, w, b, add=add, epi=ops.EPI_RELU)
    copies = {}
    n0 = ops.launch_count()
    got = ops.conv_forward(spec, x, w, b, add=add, epi=ops.EPI_RELU, bf16_copy=copies)
    assert ops.launch_count() == n0 + 1, "the copy must come from the convolution's own epilogue, not from a second kernel"
    c = copies[got.data_ptr()]
    assert c.dtype == torch.bfloat16 and c.shape == got.shape and c.stride() == got.stride()
    assert torch.equal(got, plain)
    # both are roundings of one fp32 value v: |fp16(v) - bf16(v)| <= 2^-9 |v| + 2^-12 |v|; and the copy is zero exactly where ReLU cut
    d = (c.float() - got.float()).abs()
    assert (d <= got.float().abs() * (2.0 ** -8) + 1e-7).all(), d.max().item()
    assert torch.equal(c == 0, got == 0) or ((c == 0) != (got == 0)).sum().item() <= 2      # (values below the fp16 subnormal range)


def test_tc_backward_weight_scratch_flush_matches_direct_flush(monkeypatch):
    """The two flushes of the backward-weight kernel -- scalar atomics straight into the parameter gradient, or vector reductions into the
    [tap][X channel][DY channel] scratch + ONE scatter launch for several layers (ops.WgradScratch) -- accumulate the same sums: a conv, a transposed
    conv and a zero-padded layer share one scratch and one flush here; the scratch is zero again afterwards."""
    from spaa_b200 import ops
    B = 2
    layers = [(ops.ConvSpec("conv", 64, 128, 3, 1, 1), 64, 128, 12, 20, 0, None), (ops.ConvSpec("convT", 64, 32, 2, 2, 0), 64, 32, 9, 11, 0, None),
              (ops.ConvSpec("conv", 6, 32, 3, 2, 1), 16, 32, 14, 18, 3, 6)]
    tensors = []
    for n, (spec, cx, cy, H, W, xoff, real) in enumerate(layers):
        x = torch.zeros(B, cx, H, W)
        x[:, xoff:xoff + (real or cx)] = synth.randn(60 + n, "sf.x", (B, real or cx, H, W))
        Ho, Wo = spec.out_hw(H, W)
        dy = synth.randn(70 + n, "sf.dy", (B, cy, Ho, Wo))
        tensors.append((cl(x), cl(dy)))
    monkeypatch.setattr(ops, "WGRAD_SCRATCH", False)
    direct = []
    for (spec, *_r, xoff, real), (x, dy) in zip(layers, tensors):
        dw = torch.full(spec.weight_shape(), 0.25, device="cuda:0")
        ops.conv_backward_weight(spec, x, dy, dw, None, x_offset=xoff)
        direct.append(dw)
    monkeypatch.setattr(ops, "WGRAD_SCRATCH", True)
    sc = ops.wgrad_scratch("cuda:0")
    got = []
    n0 = ops.launch_count()
    for (spec, *_r, xoff, real), (x, dy) in zip(layers, tensors):
        dw = torch.full(spec.weight_shape(), 0.25, device="cuda:0")
        ops.conv_backward_weight(spec, x, dy, dw, None, x_offset=xoff, scratch=sc)
        got.append(dw)
    assert all((g == 0.25).all() for g in got), "nothing reaches the parameter gradients before the flush"
    sc.flush()
    assert ops.launch_count() == n0 + len(layers) + 1, "one kernel per layer + ONE scatter launch"
    for a, b in zip(direct, got):
        tol = 1e-5 * max(1.0, a.abs().max().item())                 # same products, different summation order of the fp32 reductions
        assert (a - b).abs().max().item() <= tol, (a - b).abs().max().item()
    assert sc.buf is not None and not sc.buf.any().item(), "the scatter launch leaves the scratch zero-filled for the next pass"
