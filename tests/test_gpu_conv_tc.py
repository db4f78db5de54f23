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


def cl(t):
    return t.to("cuda:0").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def close(a, b, atol, rtol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs() - rtol * b.abs()
    assert err.max().item() <= atol, f"{what}: max abs err {(a - b).abs().max().item():.3e}"


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_tc_conv_forward_and_backward_data(case):
    from spaa_b200 import ops
    from spaa_b200._lib import lib
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
