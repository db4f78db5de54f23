"""GPU parity tests of the individual kernels (through the C ABI) against the CPU oracle / golden fixtures."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import spaa_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.asarray(a))


def maxerr(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def close(a, b, atol, rtol=0.0, what=""):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(torch.isnan(a), torch.isnan(b)), what + ": NaN pattern"
    ok = ~torch.isnan(a)
    err = ((a - b).abs() - rtol * b.abs())[ok]
    assert err.numel() == 0 or err.max().item() <= atol, f"{what}: max abs err {(a - b).abs()[ok].max().item():.3e} (atol {atol}, rtol {rtol})"


# ---------------------------------------------------------------------------------------------------------
# colour
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["rand", "edge"])
def test_colour_vs_reference_golden(golden, tag):
    from spaa_b200 import ops
    g = golden("colour")
    x, y = T(g[tag + "_x"]).to(dev()), T(g[tag + "_y"]).to(dev())
    lx, ly = ops.rgb2lab(x), ops.rgb2lab(y)
    close(lx, g[tag + "_labx"], 1e-4, 1e-6, "labx")
    close(ly, g[tag + "_laby"], 1e-4, 1e-6, "laby")
    # feed the reference's own Lab so dE is compared on identical inputs
    rlx, rly = T(g[tag + "_labx"]).to(dev()), T(g[tag + "_laby"]).to(dev())
    de = ops.de2000(rlx, rly)
    close(de, g[tag + "_de"], 2e-5, 1e-5, "de")
    cot = T(g[tag + "_cot"]).to(dev())
    d1, d2 = ops.de2000_bwd(lx, ly, cot)
    gx, gy = ops.rgb2lab_bwd(x, d1), ops.rgb2lab_bwd(y, d2)
    # The gradient is ill-conditioned in fp32 (hue of low-chroma colours, 1/dE): the reference's own fp32 result is
    # several 1e-3 away from a float64 evaluation.  Requirement: the kernel is as close to the float64 truth as the
    # reference is (factor 3 + 1e-4), and within 1e-2 + 1e-4 rel of the reference itself on well-conditioned pixels.
    xd_, yd_ = T(g[tag + "_x"]).double().requires_grad_(True), T(g[tag + "_y"]).double().requires_grad_(True)
    de64 = O.de2000_variant(O.srgb_to_lab(xd_), O.srgb_to_lab(yd_))
    tx, ty_ = torch.autograd.grad((de64 * T(g[tag + "_cot"]).double()).sum(), (xd_, yd_))
    chroma = torch.minimum(T(g[tag + "_labx"])[:, 1:].norm(dim=1), T(g[tag + "_laby"])[:, 1:].norm(dim=1))
    ok = (chroma > 0.5).unsqueeze(1).expand_as(gx.cpu())
    for got, ref, truth, nm in ((gx.cpu(), T(g[tag + "_gx"]), tx, "gx"), (gy.cpu(), T(g[tag + "_gy"]), ty_, "gy")):
        close(got[ok], ref[ok], 1e-2, 1e-4, nm)
        close(got[~ok], ref[~ok], 1e-2, 0.5, nm + " near-neutral")
        fin = torch.isfinite(truth) & torch.isfinite(ref.double()) & ok
        e_ref = (ref.double() - truth).abs()[fin].max().item()
        e_got = (got.double() - truth).abs()[fin].max().item()
        assert e_got <= 3 * e_ref + 1e-4, f"{nm}: kernel is {e_got:.2e} from the float64 truth, the reference {e_ref:.2e}"


@pytest.mark.parametrize("cam_is_lab2,de_weighting", [(False, False), (True, True)])
def test_fused_colour_loss(cam_is_lab2, de_weighting):
    from spaa_b200 import ops
    B, H, W = 3, 37, 53
    scene = synth.textured(5, "cl.scene", (1, 3, H, W))
    cam = (scene + synth.randn(6, "cl.cam", (B, 3, H, W), 0.05)).clamp(0, 1)
    cam[0, :, :5] = scene[0, :, :5]                       # identical pixels: dE = 0, L2 = 0, zero gradient
    c_de, c_l2 = 0.7 / (H * W), 1.3 / (H * W)
    x = cam.clone().requires_grad_(True)
    la, lb = O.srgb_to_lab(x), O.srgb_to_lab(scene.expand(B, -1, -1, -1))
    de = O.de2000_variant(lb, la) if cam_is_lab2 else O.de2000_variant(la, lb)
    l2 = torch.norm(x - scene, dim=1)
    obj = c_de * ((0.5 * de ** 2) if de_weighting else de).sum() + c_l2 * l2.sum()
    gref, = torch.autograd.grad(obj, x)
    ref_lab = ops.rgb2lab(scene.to(dev()))
    stats, grad = ops.color_loss(cam.to(dev()), scene.to(dev()), ref_lab, cam_is_lab2=cam_is_lab2, de_weighting=de_weighting,
                                 c_de=c_de, c_l2=c_l2)
    close(stats[:, 0], de.sum((1, 2)), 1e-2, 1e-5, "sum dE")
    close(stats[:, 1], l2.sum((1, 2)), 1e-3, 1e-5, "sum L2")
    close(stats[:, 2], (de ** 2).sum((1, 2)), 1e-1, 1e-5, "sum dE^2")
    close(grad, gref, 2e-6, 2e-3, "grad")
    assert torch.isfinite(grad).all()
    # launching again with the same workspace gives the same statistics (self-resetting counters)
    stats2, _ = ops.color_loss(cam.to(dev()), scene.to(dev()), ref_lab, cam_is_lab2=cam_is_lab2, de_weighting=de_weighting,
                               c_de=c_de, c_l2=c_l2, want_grad=False)
    assert torch.equal(stats, stats2)
    # the hardware-approximation arithmetic of the 16-bit modes against the exact kernel: statistics to 2e-5 relative, the gradient to 1e-3 of
    # its largest entry plus 1 % (the hue terms amplify a 1e-6 relative error of Lab near neutral colours)
    ref_lab_f = ops.rgb2lab(scene.to(dev()), fast=True)          # same arithmetic on both sides: equal pixels -> equal Lab -> dE = 0 exactly
    close(ref_lab_f, ref_lab, 2e-4, 2e-6, "fast-arithmetic Lab")
    stats_f, grad_f = ops.color_loss(cam.to(dev()), scene.to(dev()), ref_lab_f, cam_is_lab2=cam_is_lab2, de_weighting=de_weighting,
                                     c_de=c_de, c_l2=c_l2, fast=True)
    close(stats_f[:, :3], stats[:, :3], 1e-3, 2e-5, "fast-arithmetic statistics")
    assert torch.isfinite(grad_f).all()
    assert torch.equal(grad_f[0, :, :5], torch.zeros_like(grad_f[0, :, :5])), "identical pixels keep an exactly zero gradient"
    close(grad_f, grad, 1e-3 * grad.abs().max().item(), 1e-2, "fast-arithmetic gradient")
    rel = ((grad_f - grad).double().norm() / grad.double().norm()).item()
    assert rel <= 1e-4, rel


# ---------------------------------------------------------------------------------------------------------
# warping
# ---------------------------------------------------------------------------------------------------------

def _planar(grid_bhw2):
    return grid_bhw2[0].permute(2, 0, 1).contiguous()


def test_tps_and_coarse_grid_forward_backward():
    from spaa_b200 import ops
    P = synth.warping_params(21, theta_scale=0.02)
    aff, theta, ctrl = P["warping_net.affine_mat"], P["warping_net.theta"], P["warping_net.ctrl_pts"]
    for (ih, iw), (oh, ow) in (((20, 24), (12, 16)), ((256, 256), (240, 320))):
        ref = _planar(O.tps_sampling_grid(theta, ctrl, oh, ow))
        close(ops.tps_grid(theta.to(dev()), ctrl.to(dev()), oh, ow), ref, 2e-6, 0, "tps grid")
        a = aff.clone().requires_grad_(True)
        t = theta.clone().requires_grad_(True)
        # coarse grid = the TPS grid sampling the affine grid (no refinement, no clamp)
        g = F.grid_sample(O.affine_base_grid(a, ih, iw).permute(0, 3, 1, 2), O.tps_sampling_grid(t, ctrl, oh, ow), align_corners=True)
        got = ops.coarse_grid(aff.to(dev()), theta.to(dev()), ctrl.to(dev()), (ih, iw), (oh, ow))
        close(got, g[0], 1e-4, 0, "coarse grid")     # the fp32 reference itself is 3e-5 from a float64 evaluation at 240x320
        cot = synth.randn(3, f"cg.cot{oh}", (2, oh, ow))
        ga, gt = torch.autograd.grad((g[0] * cot).sum(), (a, t))
        da, dt = ops.coarse_grid_bwd(aff.to(dev()), theta.to(dev()), ctrl.to(dev()), (ih, iw), (oh, ow), cot.to(dev()))
        close(da, ga, 2e-4 * max(1.0, ga.abs().max().item()), 1e-3, "d affine")
        close(dt, gt, 2e-4 * max(1.0, gt.abs().max().item()), 1e-3, "d theta")


def test_grid_finish():
    from spaa_b200 import ops
    c = synth.randn(1, "gf.c", (2, 9, 11), 0.8)
    r = synth.randn(2, "gf.r", (2, 9, 11), 0.5)
    c[0, 0, 0], r[0, 0, 0] = 0.5, 0.5          # exactly on the bound: clamp backward passes the gradient
    cc = c.clone().requires_grad_(True)
    f = torch.clamp(r + cc, -1, 1)
    cot = synth.randn(3, "gf.cot", (2, 9, 11))
    gref, = torch.autograd.grad((f * cot).sum(), cc)
    close(ops.grid_finish(c.to(dev()), r.to(dev())), f, 0, 0, "fine")
    close(ops.grid_finish_bwd(c.to(dev()), r.to(dev()), cot.to(dev())), gref, 0, 0, "dfine")


@pytest.mark.parametrize("shared_grid", [True, False])
def test_grid_sample_forward_backward(shared_grid):
    from spaa_b200 import ops
    B, C, Hi, Wi, H, W = 3, 3, 20, 24, 15, 18
    img = synth.randn(7, "gs.img", (B, C, Hi, Wi), 0.6) + 0.5          # some values outside [0,1]
    gb = 1 if shared_grid else B
    grid = synth.rand(8, "gs.grid", (gb, H, W, 2)) * 2.3 - 1.15        # some samples fall outside the image
    mask = (synth.rand(9, "gs.mask", (H, W)) > 0.2).float()
    rough = synth.rand(10, "gs.rough", (B, C, H, W))
    x = img.clone().requires_grad_(True)
    gq = grid.clone().requires_grad_(True)
    y = F.grid_sample(torch.clamp(x, 0, 1), gq.expand(B, -1, -1, -1), align_corners=True) * mask
    y2 = y * rough
    c1, c2 = synth.randn(11, "gs.c1", y.shape), synth.randn(12, "gs.c2", y.shape)
    gx, gg = torch.autograd.grad((y * c1).sum() + (y2 * c2).sum(), (x, gq))
    gp = grid.permute(0, 3, 1, 2).contiguous()
    gp = gp[0] if shared_grid else gp
    wide = torch.zeros(B, 6, H, W, device=dev())
    out = ops.grid_sample(img.to(dev()), gp.to(dev()), clamp01=True, mask=mask.flatten().to(dev()), rough=rough.to(dev()), out2=wide[:, 3:])
    close(out, y, 2e-6, 0, "out")
    close(wide[:, 3:], y2, 2e-6, 0, "out2")
    dimg = ops.grid_sample_bwd_input(c1.to(dev()), gp.to(dev()), (Hi, Wi), mask=mask.flatten().to(dev()), dout2=c2.to(dev()), rough=rough.to(dev()))
    inside = ((img >= 0) & (img <= 1)).float()
    close(dimg.cpu() * inside, gx, 1e-5, 1e-5, "dimg")
    if shared_grid:
        # gather form through the per-attack CSR adjoint map: same gradient, no atomics; fused squared norm of the clamp-masked gradient
        adj = ops.WarpAdjoint(gp.to(dev()), (Hi, Wi), mask.flatten().to(dev()))
        sq = torch.empty(B, device=dev())
        dimg2 = ops.grid_sample_bwd_gather(adj, c1.to(dev()), dout2=c2.to(dev()), rough=rough.to(dev()), sq=sq, x_for_clamp=img.to(dev()))
        close(dimg2.cpu() * inside, gx, 1e-5, 1e-5, "dimg (gather)")
        close(sq, (gx.double() ** 2).flatten(1).sum(1), 1e-5, 1e-5, "fused squared norm")
        again = ops.grid_sample_bwd_gather(adj, c1.to(dev()), dout2=c2.to(dev()), rough=rough.to(dev()))
        assert torch.equal(again, dimg2), "the gather adjoint must be deterministic"
        # the tiled (shared-memory staged) kernel and the plain one run the same entries in the same order: bit-identical gradient and norm
        assert ops.GATHER_TILED and adj.max_region > 0
        ops.GATHER_TILED = False
        try:
            sq_p = torch.empty(B, device=dev())
            plain = ops.grid_sample_bwd_gather(adj, c1.to(dev()), dout2=c2.to(dev()), rough=rough.to(dev()), sq=sq_p, x_for_clamp=img.to(dev()))
        finally:
            ops.GATHER_TILED = True
        assert torch.equal(plain, dimg2) and torch.allclose(sq_p, sq, rtol=1e-6, atol=0), "tiled and plain gather differ"
        no_mask = ops.WarpAdjoint(gp.to(dev()), (Hi, Wi))
        d3 = ops.grid_sample_bwd_gather(no_mask, c1.to(dev()))
        y3 = F.grid_sample(x, gq.expand(B, -1, -1, -1), align_corners=True)
        g3, = torch.autograd.grad((y3 * c1).sum(), x)
        close(d3, g3, 1e-5, 1e-5, "dimg (gather, no mask / rough / clamp)")
    dgrid = ops.grid_sample_bwd_grid(c1.to(dev()), img.to(dev()), gp.to(dev()), clamp01=True, mask=mask.flatten().to(dev()),
                                     dout2=c2.to(dev()), rough=rough.to(dev()))
    ref = gg.permute(0, 3, 1, 2)
    close(dgrid, ref[0] if shared_grid else ref, 2e-4, 1e-5, "dgrid")


# ---------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------

CONV_CASES = [
    # kind, cin, cout, k, stride, pad, outpad, H, W
    ("conv", 3, 32, 3, 2, 1, 0, 24, 32), ("conv", 32, 64, 3, 2, 1, 0, 13, 17), ("conv", 64, 128, 3, 1, 1, 0, 9, 12),
    ("conv", 32, 3, 3, 1, 1, 0, 24, 32), ("conv", 3, 3, 1, 1, 0, 0, 10, 11), ("conv", 32, 64, 1, 1, 0, 0, 12, 16),
    ("conv", 6, 32, 3, 2, 1, 0, 24, 32), ("conv", 2, 32, 3, 2, 1, 0, 24, 32), ("conv", 20, 24, 3, 1, 1, 0, 7, 9),
    ("convT", 128, 64, 3, 2, 1, 1, 6, 8), ("convT", 64, 32, 2, 2, 0, 0, 6, 8), ("convT", 32, 2, 2, 2, 0, 0, 12, 16),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_forward_backward(case):
    from spaa_b200 import ops
    kind, cin, cout, k, stride, pad, outpad, H, W = case
    B = 2
    spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
    x = synth.randn(31, "cv.x", (B, cin, H, W)).double().requires_grad_(True)
    w = synth.randn(32, "cv.w", spec.weight_shape(), (2.0 / (cin * k * k)) ** 0.5).double().requires_grad_(True)
    b = synth.randn(33, "cv.b", (cout,), 0.1).double().requires_grad_(True)
    if kind == "conv":
        pre = F.conv2d(x, w, b, stride, pad)
    else:
        pre = F.conv_transpose2d(x, w, b, stride, pad, outpad)
    Ho, Wo = pre.shape[-2:]
    assert (Ho, Wo) == spec.out_hw(H, W)
    add = synth.randn(34, "cv.add", pre.shape, 0.5).double()
    y = torch.clamp(F.relu(pre + add), max=1)
    xd, wd, bd, addd = x.detach().float().to(dev()), w.detach().float().to(dev()), b.detach().float().to(dev()), add.float().to(dev())
    got = ops.conv_forward(spec, xd, wd, bd, add=addd, epi=ops.EPI_RELU | ops.EPI_CLAMP_MAX1)
    close(got, y, 2e-5, 1e-5, "forward")
    # leaky + add-after-activation epilogue (the refinement net's last layer)
    y2 = F.leaky_relu(pre, 0.1) + add
    close(ops.conv_forward(spec, xd, wd, bd, add=addd, epi=ops.EPI_LEAKY01 | ops.EPI_ADD_AFTER_ACT), y2, 2e-5, 1e-5, "leaky")
    # backward of the plain pre-activation
    cot = synth.randn(35, "cv.cot", pre.shape).double()
    gx, gw, gb = torch.autograd.grad((pre * cot).sum(), (x, w, b))
    cotd = cot.float().to(dev())
    m = synth.randn(36, "cv.m", x.shape).double()                       # stands for the producer's activation
    m2 = synth.randn(37, "cv.m2", x.shape).double()
    extra = synth.randn(38, "cv.extra", x.shape).double()
    out2 = torch.empty(x.shape, device=dev())
    dx = ops.conv_backward_data(spec, cotd, wd, (H, W), add=extra.float().to(dev()), mask=m.float().to(dev()), mask_mode=ops.MASK_POS,
                                mask2=m2.float().to(dev()), out2=out2)
    ref = (gx + extra) * (m > 0)
    close(dx, ref, 3e-5, 1e-5, "bwd data")
    close(out2, ref * (m2 > 0), 3e-5, 1e-5, "bwd data out2")
    if cin >= 6:      # channel-sliced weight view -> gradient of only some input channels
        wv = wd[:, 3:6] if kind == "conv" else wd[3:6]
        close(ops.conv_backward_data(spec, cotd, wv, (H, W)), gx[:, 3:6], 3e-5, 1e-5, "bwd data (slice)")
    dw, db = torch.zeros_like(wd), torch.zeros_like(bd)
    ops.conv_backward_weight(spec, xd, cotd, dw, db)
    scale = max(1.0, gw.abs().max().item())
    close(dw, gw, 2e-5 * scale, 1e-5, "bwd weight")
    close(db, gb, 2e-5 * max(1.0, gb.abs().max().item()), 1e-5, "bwd bias")


def test_conv_channels_last_and_bf16():
    from spaa_b200 import ops
    spec = ops.ConvSpec("conv", 16, 40, 3, 1, 1)
    x = synth.randn(41, "cl.x", (2, 16, 9, 10))
    w = synth.randn(42, "cl.w", spec.weight_shape(), 0.1)
    b = synth.randn(43, "cl.b", (40,), 0.1)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), 1, 1))
    xcl = x.to(dev()).contiguous(memory_format=torch.channels_last)
    out = torch.empty((2, 40, 9, 10), device=dev()).contiguous(memory_format=torch.channels_last)
    ops.conv_forward(spec, xcl, w.to(dev()), b.to(dev()), out=out, epi=ops.EPI_RELU)
    close(out, ref, 2e-5, 1e-5, "channels_last")
    xb = xcl.to(torch.bfloat16)
    refb = F.relu(F.conv2d(xb.float().cpu().double(), w.double(), b.double(), 1, 1))
    outb = ops.conv_forward(spec, xb, w.to(dev()), b.to(dev()), epi=ops.EPI_RELU, out_dtype=torch.bfloat16)
    close(outb.float(), refb, 2e-2, 1e-2, "bf16 storage")


# ---------------------------------------------------------------------------------------------------------
# SSIM + L1 loss, Adam
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("hw", [(40, 48), (33, 70), (13, 9)])
def test_ssim_l1_fused(hw):
    from spaa_b200 import ops
    B, C = 2, 3
    tgt = synth.textured(51, "ss.t", (B, C, *hw))
    pred = (tgt + synth.randn(52, "ss.p", (B, C, *hw), 0.08)).clamp(0, 1)
    p = pred.clone().double().requires_grad_(True)
    m = O.ssim_map(p, tgt.double())
    l1, l2 = (p - tgt.double()).abs().mean(), ((p - tgt.double()) ** 2).mean()
    loss = 0.9 * l1 + 0.3 * l2 + 1.1 * (1 - m.mean())
    gref, = torch.autograd.grad(loss, p)
    sums, grad, smap = ops.ssim_l1(pred.to(dev()), tgt.to(dev()), 0.9, 0.3, 1.1, want_map=True)
    n = pred.numel()
    close(smap, m, 2e-4, 0, "ssim map")      # fp32 cancellation in E[x^2]-mu^2; truth is float64
    close(sums[0] / n, l1, 1e-6, 1e-5, "l1")
    close(sums[1] / n, l2, 1e-6, 1e-5, "l2")
    close(sums[2] / n, m.mean(), 1e-5, 0, "ssim mean")
    close(grad, gref, 2e-7, 2e-3, "grad")
    # per-pixel cotangent map (size_average=False / weights / mask branches)
    cot = synth.randn(53, "ss.cot", pred.shape, 1.0 / n)
    gref2, = torch.autograd.grad((O.ssim_map(p, tgt.double()) * cot.double()).sum(), p)
    _, grad2, _ = ops.ssim_l1(pred.to(dev()), tgt.to(dev()), 0, 0, 0, cot_map=cot.to(dev()))
    close(grad2, gref2, 2e-7, 2e-3, "grad (cot map)")


def test_adam_flat():
    from spaa_b200 import ops
    n = 1000
    p0, g = synth.randn(61, "ad.p", (n,)), synth.randn(62, "ad.g", (n,), 0.1)
    seg_end = torch.tensor([300, 650, n], dtype=torch.int64)
    lrs, wds = [1e-2, 5e-3, 1e-3], [0.0, 0.0, 1e-4]
    pr, mr, vr = p0.clone(), torch.zeros(n), torch.zeros(n)
    pd, md, vd = p0.to(dev()), torch.zeros(n, device=dev()), torch.zeros(n, device=dev())
    for step in (1, 2, 3):
        gs = g * step
        lo = 0
        for e, lr, wd in zip(seg_end.tolist(), lrs, wds):
            O.adam_step(pr[lo:e], gs[lo:e], mr[lo:e], vr[lo:e], step, lr, wd)
            lo = e
        ops.adam_step(pd, gs.to(dev()), md, vd, seg_end.to(dev()), torch.tensor(lrs, device=dev()), torch.tensor(wds, device=dev()), step)
    close(pd, pr, 1e-6, 1e-5, "adam params")


# ---------------------------------------------------------------------------------------------------------
# attack-loop kernels
# ---------------------------------------------------------------------------------------------------------

def test_row_norm_step_copy_select():
    from spaa_b200 import ops
    B, n = 5, 3 * 17 * 19
    x = synth.randn(71, "at.x", (B, n), 0.6) + 0.5
    g = synth.randn(72, "at.g", (B, n))
    sel = torch.tensor([1, 0, 1, 0, 0], dtype=torch.uint8)
    inside = ((x >= 0) & (x <= 1)).float()
    gm = g * inside
    sq = torch.empty(B, device=dev())
    ops.row_sqnorm(g.to(dev()), sq, x.to(dev()))
    close(sq, (gm ** 2).sum(1), 1e-2, 1e-5, "sqnorm")
    step2 = torch.tensor([-2.0, -1.0], device=dev())
    xd = x.to(dev()).clone()
    best = torch.zeros(B, n, device=dev())
    copy_sel = torch.tensor([0, 1, 1, 0, 0], dtype=torch.uint8)
    ops.row_normalized_step(xd, g.to(dev()), sq, step2, sel.to(dev()), use_clamp_mask=True, copy_dst=best, copy_sel=copy_sel.to(dev()))
    stepv = torch.where(sel.bool(), torch.tensor(-1.0), torch.tensor(-2.0)).view(B, 1)
    ref = x + stepv * gm / gm.norm(dim=1, keepdim=True)
    close(xd, ref, 1e-6, 1e-6, "step")
    close(best, ref * copy_sel.view(B, 1), 1e-6, 1e-6, "copy")
    # zero step leaves rows untouched; sum_out written for all rows
    xd2 = x.to(dev()).clone()
    base = synth.rand(73, "at.base", (1, n)).to(dev())
    so = torch.empty(B, n, device=dev())
    ops.row_normalized_step(xd2, g.to(dev()), sq, torch.tensor([0.0, 0.5], device=dev()), sel.to(dev()), base=base, sum_out=so)
    ref2 = torch.where(sel.bool().view(B, 1), x + 0.5 * g / gm.norm(dim=1, keepdim=True), x)
    close(xd2, ref2, 1e-6, 1e-6, "step (zero rows)")
    close(so, base.cpu() + ref2, 1e-6, 1e-6, "sum_out")
    dst = torch.zeros(B, n, device=dev())
    ops.masked_copy_rows(dst, x.to(dev()), sel.to(dev()))
    close(dst, x * sel.view(B, 1), 0, 0, "masked copy")
    act = synth.rand(74, "at.act", (B, n)) * 1.4 - 0.2
    act = torch.where(act > 1, torch.ones(()), torch.where(act < 0, torch.zeros(()), act))
    out = torch.empty(B, n, device=dev())
    ops.select_cotangent(g.to(dev()), x.to(dev()), sel.to(dev()), act.to(dev()), ops.MASK_OPEN01, out)
    refc = torch.where(sel.bool().view(B, 1), x, g) * ((act > 0) & (act < 1))
    close(out, refc, 0, 0, "select cotangent")


def test_percal_project_and_chan_l2():
    from spaa_b200 import ops
    B, H, W = 3, 11, 13
    base = synth.rand(81, "pp.base", (1, 3, H, W))
    delta = synth.randn(82, "pp.delta", (B, 3, H, W), 0.3)
    d = (base + delta).clamp(0, 1) - base
    xs = base + d
    xq = torch.round(xs * 255) / 255
    l2 = torch.norm(d, dim=1).sum((1, 2))
    dd, xqd, xsd, l2d = delta.to(dev()).clone(), torch.empty(B, 3, H, W, device=dev()), torch.empty(B, 3, H, W, device=dev()), torch.empty(B, device=dev())
    ops.percal_project(base.to(dev()), dd, xqd, xsd, l2d)
    close(dd, d, 0, 0, "delta"); close(xsd, xs, 0, 0, "xsum"); close(xqd, xq, 0, 0, "xq"); close(l2d, l2, 1e-3, 1e-5, "l2sum")
    x = synth.randn(83, "cl2.x", (B, 3, H, W), 0.5) + 0.5
    ref = 0.5 * torch.ones(1, 3, H, W)
    xr = x.clone().requires_grad_(True)
    s = torch.norm(ref - xr, dim=1).sum((1, 2))
    gref, = torch.autograd.grad(0.37 * s.sum(), xr)
    g0 = synth.randn(84, "cl2.g", (B, 3, H, W))
    sel = torch.tensor([1, 0, 1], dtype=torch.uint8)
    sums, gd = torch.empty(B, device=dev()), g0.to(dev()).clone()
    ops.chan_l2(x.to(dev()), ref.to(dev()), sums, c=0.37, sel=sel.to(dev()), apply_clamp_mask=True, grad=gd)
    close(sums, s, 1e-3, 1e-5, "chan l2 sums")
    close(gd, g0 * ((x >= 0) & (x <= 1)) + gref * sel.view(B, 1, 1, 1), 1e-6, 1e-5, "chan l2 grad")


def test_attack_masks():
    from spaa_b200 import ops
    B, ncls, hw = 6, 1000, 100
    logits = synth.randn(91, "am.l", (B, ncls), 3.0)
    target = torch.tensor([3, 7, 11, 500, 999, 0])
    logits[0, 3] = 40.0; logits[1, 7] = 12.0; logits[2, 5] = 50.0; logits[3, 500] = 45.0
    stats = torch.tensor([[300., 3.0, 0, 0], [300., 3.0, 0, 0], [300., 3.0, 0, 0], [300., 0.1, 0, 0], [100., 9.0, 0, 0], [1., 4., 0, 0]])
    p = F.softmax(logits, 1)
    pm, am = p.max(1)
    caml2, camde = stats[:, 1] / hw, stats[:, 0] / hw
    col = 1.0 * caml2 + 1.0 * camde
    for targeted in (True, False):
        best = torch.tensor([1e6, 1e6, 1e6, 1e6, 0.5, 1e6])
        succ = (am == target) if targeted else (am != target)
        use = succ & (caml2 * 255 > 5) & ((pm > 0.9) if targeted else torch.ones(B, dtype=torch.bool))
        better = use & (col < best)
        bexp = torch.where(better, col, best)
        u, s, bt = (torch.empty(B, dtype=torch.uint8, device=dev()) for _ in range(3))
        cl, bd = torch.empty(B, device=dev()), best.to(dev())
        ops.attack_masks(logits.to(dev()), target.to(dev()), targeted, stats.to(dev()), None, hw, 0, 0.0, 1.0, 1.0, 5.0, 0.9, u, s, bt, cl, bd)
        assert torch.equal(u.cpu().bool(), use) and torch.equal(s.cpu().bool(), succ) and torch.equal(bt.cpu().bool(), better)
        close(cl, col, 1e-6, 1e-6, "col loss"); close(bd, bexp, 1e-6, 1e-6, "best col")
    # PerC-AL margin mode
    l2sum = torch.tensor([3.0, 3.0, 0.1, 3.0, 3.0, 3.0])
    stats[:, 2] = torch.tensor([4.0, 9.0, 16.0, 25.0, 36.0, 49.0])
    real = logits.gather(1, target.view(-1, 1)).squeeze(1)
    other = logits.scatter(1, target.view(-1, 1), -float("inf")).max(1)[0]
    isadv = (real - other) <= -40
    use = isadv & (l2sum / hw * 255 > 5)
    ia, u, bt = (torch.empty(B, dtype=torch.uint8, device=dev()) for _ in range(3))
    dis, bd = torch.empty(B, device=dev()), torch.full((B,), 1e5, device=dev())
    ops.percal_masks(logits.to(dev()), target.to(dev()), 2, 40.0, l2sum.to(dev()), hw, 5.0, 0.9, stats.to(dev()), ia, u, bt, dis, bd)
    assert torch.equal(ia.cpu().bool(), isadv) and torch.equal(u.cpu().bool(), use)
    close(dis, stats[:, 2].sqrt(), 1e-6, 1e-6, "dis")


@pytest.mark.parametrize("hw,crop,insz", [((240, 320), (240, 240), (224, 224)), ((240, 320), (240, 240), (299, 299)), ((24, 32), (24, 24), (20, 20)),
                                          ((30, 41), (27, 33), (11, 40))])
@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_classifier_preprocess_vs_oracle(hw, crop, insz, channels_last):
    """classifier.py:55-59 (centre crop, area resize, normalise) as one kernel, and its adjoint, against the oracle + autograd."""
    from spaa_b200 import ops
    from spaa_b200.classifier import preprocess_fused
    B = 3
    im = synth.rand(51, "pre.im", (B, 3, *hw))
    ref_in = im.clone().double().requires_grad_(True)
    ref = O.classifier_preprocess(ref_in, crop, insz)
    cot = synth.randn(52, "pre.cot", ref.shape).double()
    gref, = torch.autograd.grad((ref * cot).sum(), ref_in)
    x = im.to(dev()).requires_grad_(True)
    n0 = ops.launch_count()
    y = preprocess_fused(x, crop, insz, channels_last)
    assert ops.launch_count() == n0 + 1, "the fused kernel was not used"
    assert y.shape == ref.shape and (y.is_contiguous(memory_format=torch.channels_last) if channels_last else y.is_contiguous())
    close(y, ref, 2e-6, 1e-6, "preprocess forward")
    c = cot.float().to(dev())
    if channels_last:
        c = c.contiguous(memory_format=torch.channels_last)
    (y * c).sum().backward()
    close(x.grad, gref, 1e-6, 1e-5, "preprocess backward")


# ---------------------------------------------------------------------------------------------------------
# fused ReLU + max-pooling (the external classifier's first stage in the attack engines' private copy)
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("k,s,p,hw,C,relu", [(3, 2, 1, (112, 112), 64, True), (2, 2, 0, (56, 40), 128, True), (3, 2, 0, (147, 147), 64, False),
                                             (3, 2, 1, (9, 7), 4, True), (2, 2, 0, (7, 9), 8, True), (3, 1, 1, (6, 5), 12, False), (3, 3, 1, (8, 8), 4, True)])
def test_fused_relu_maxpool_vs_torch(k, s, p, hw, C, relu):
    """Values are exact (max-pooling selects); gradients equal torch's relu + max_pool2d autograd (same arg-max convention: first maximum in
    row-major window order; ties at 0 carry no gradient through the ReLU) up to the summation order of <= 4 overlapping windows.  Inputs are
    quantised to multiples of 0.5 with a third of them zero, so ties -- positive ones and the all-zero windows a ReLU produces -- are everywhere."""
    from spaa_b200.classifier import FusedReLUMaxPool2d
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, C, *hw, generator=g)
    x = torch.round(x * 2) / 2
    x[torch.rand(x.shape, generator=g) < 0.3] = 0.0
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(F.relu(xr) if relu else xr, k, s, p)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xd = x.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bias = (torch.round(torch.randn(C, generator=g) * 2) / 2) if (relu and C >= 8) else None     # the stripped bias of the convolution in front
    if bias is not None:
        xr = x.clone().requires_grad_(True)
        yr = F.max_pool2d(F.relu(xr + bias.view(1, -1, 1, 1)), k, s, p)
        yr.backward(dy)
    m = FusedReLUMaxPool2d(k, s, p, with_relu=relu, bias=bias).to(dev())
    from spaa_b200 import ops
    n0 = ops.launch_count()
    y = m(xd)
    assert ops.launch_count() == n0 + 1, "the fused kernel did not run"
    assert y.is_contiguous(memory_format=torch.channels_last) or min(y.shape) == 1
    y.backward(dy.to(dev()))
    assert ops.launch_count() == n0 + 2
    assert torch.equal(y.detach().cpu(), yr.detach()), "pooled values must be exact"
    close(xd.grad, xr.grad, 1e-6, 1e-6, "fused relu+maxpool backward")
    # NCHW-contiguous gradients are accepted too (converted), and non-channels_last inputs run the stock ops of the pair
    xd2 = x.to(dev()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    m(xd2).backward(dy.to(dev()).contiguous())
    close(xd2.grad, xr.grad, 1e-6, 1e-6, "fused relu+maxpool backward (NCHW cotangent)")
    if C > 1 and hw[0] > 1:
        assert torch.equal(m(x.to(dev())).cpu(), yr.detach())


@pytest.mark.parametrize("shape,with_bias,with_res,relu", [((3, 64, 56, 56), True, True, True), ((2, 128, 7, 9), True, False, True),
                                                           ((1, 4, 5, 3), False, True, True), ((2, 192, 35, 35), True, False, False),
                                                           ((5, 8, 1, 1), True, True, True)])
def test_fused_bias_act_vs_torch(shape, with_bias, with_res, relu):
    """relu?(x + bias + res) in one kernel: bit-identical to the stock ops applied in the same order; adjoint = threshold_backward."""
    from spaa_b200 import ops
    from spaa_b200.classifier import bias_act
    g = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=g)
    bias = torch.randn(shape[1], generator=g) if with_bias else None
    res = torch.randn(shape, generator=g) if with_res else None
    dy = torch.randn(shape, generator=g)
    xr = x.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    yr = bias_act(xr, bias, rr, relu)                      # CPU tensors: the stock-op branch
    yr.backward(dy)
    cl = lambda t: t.to(dev()).contiguous(memory_format=torch.channels_last)
    xd = cl(x).requires_grad_(True)
    rd = cl(res).requires_grad_(True) if with_res else None
    n0 = ops.launch_count()
    y = bias_act(xd, bias.to(dev()) if with_bias else None, rd, relu)
    assert ops.launch_count() == n0 + 1, "the fused kernel did not run"
    y.backward(dy.to(dev()))
    assert torch.equal(y.detach().cpu(), yr.detach())
    assert torch.equal(xd.grad.cpu(), xr.grad)
    if with_res:
        assert torch.equal(rd.grad.cpu(), rr.grad)


@pytest.mark.parametrize("name", ["resnet18", "vgg16", "inception_v3"])
def test_private_classifier_copy_pool_fusion_on_gpu(name):
    """The attack engines' private classifier copy with and without the fused ReLU + max-pooling modules (same cuDNN convolutions on both
    sides, so only the elementwise glue and the pooling differ): logits agree to fp32 rounding (the downsample branch's bias is folded into
    conv2's; window sums are re-ordered in the pooling adjoint), top-1 is identical, the input gradient agrees to 1e-4 relative, and every
    convolution of the copy is followed by exactly one of our kernels."""
    from spaa_b200.classifier import Classifier, FusedReLUMaxPool2d, fold_batchnorm, use_channels_last, device_logits
    clf = Classifier(name, dev(), [0], allow_random_init=True)
    clf.model.to(memory_format=torch.channels_last)       # (use_channels_last() declines in the exact-fp32 test configuration; the layout is what matters here)
    cl = True
    plain, fused = fold_batchnorm(clf, fuse_pool=False, fuse_bias=False, fuse_stem=False), fold_batchnorm(clf, fuse_pool=True, fuse_bias=True, fuse_stem=True)
    assert getattr(fused, "stem_s2d", False) == (name == "resnet18")
    assert any(isinstance(m, FusedReLUMaxPool2d) for m in fused.model.modules())
    assert not any(isinstance(m, FusedReLUMaxPool2d) for m in getattr(plain, "model").modules())
    x = torch.rand(4, 3, 240, 320, generator=torch.Generator().manual_seed(5)).to(dev())
    from spaa_b200 import ops
    n0 = ops.launch_count()
    with torch.no_grad():
        device_logits(fused, x, (240, 240), cl)
    n_fused_launches = ops.launch_count() - n0 - 1                                        # (- the pre-processing kernel)
    assert n_fused_launches == {"resnet18": 17, "vgg16": 13, "inception_v3": 96}[name], n_fused_launches
    outs, grads = [], []
    for c in (plain, fused):
        leaf = x.clone().requires_grad_(True)
        y = device_logits(c, leaf, (240, 240), cl)
        gq, = torch.autograd.grad(y[:, 7].sum(), leaf)
        outs.append(y.detach()); grads.append(gq)
    scale = max(outs[0].abs().max().item(), 1.0)
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-5 * scale
    assert torch.equal(outs[0].argmax(1), outs[1].argmax(1))
    rel = ((grads[0] - grads[1]).double().norm() / grads[0].double().norm()).item()
    assert rel <= 1e-4, rel


@pytest.mark.parametrize("hw,crop,insz", [((240, 320), (240, 240), (224, 224)), ((200, 260), (180, 200), (96, 128))])
def test_fused_preprocess_space_to_depth_layout(hw, crop, insz):
    """Layout 2 of the pre-processing kernel = the torch fold (classifier.s2d_fold) of its plain output, bit for bit; the adjoint through the
    folded layout equals the adjoint through the plain one; an S2DStem fed either way equals the 7x7 stride-2 convolution it stands for."""
    from spaa_b200.classifier import S2DStem, preprocess_fused, s2d_fold
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 3, *hw, generator=g).to(dev())
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    plain = preprocess_fused(x1, crop, insz, True)
    folded = preprocess_fused(x2, crop, insz, True, s2d=True)
    assert folded.shape == (3, 16, insz[0] // 2 + 3, insz[1] // 2 + 3) and folded.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(folded.detach(), s2d_fold(plain.detach()))
    cot = torch.randn(folded.shape, generator=g).to(dev())
    folded.backward(cot)
    s2d_fold(plain).backward(cot)
    close(x2.grad, x1.grad, 1e-6, 1e-6, "adjoint through the folded layout")
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3).to(dev())
    stem = S2DStem(conv).to(dev())
    with torch.no_grad():
        ref = conv(plain.detach())
        close(stem(folded.detach()), ref, 2e-5, 1e-5, "S2DStem on the kernel's folded input")
        close(stem(plain.detach()), ref, 2e-5, 1e-5, "S2DStem folding a plain input itself")


def test_channel_sum_multi_matches_per_tensor_sums():
    """Bias gradients of a whole backward pass in one launch: tensors of different widths and sizes, a zero-padded one (3 real of 16 channels),
    both 16-bit types; += semantics."""
    from spaa_b200 import ops
    dev = torch.device("cuda:0")
    shapes = [(2, 16, 24, 32, 3), (3, 32, 12, 16, 32), (1, 64, 7, 9, 64), (2, 256, 6, 8, 256), (4, 128, 5, 8, 128)]
    for dt in (torch.bfloat16, torch.float16):
        jobs, refs = [], []
        for i, (b, c, h, w, cr) in enumerate(shapes):
            x = synth.randn(40 + i, "csm.x", (b, c, h, w)).to(dev).to(dt).contiguous(memory_format=torch.channels_last)
            out = torch.full((cr,), 0.5, device=dev)
            jobs.append((x, out))
            refs.append(0.5 + x.double().sum((0, 2, 3))[:cr])
        n0 = ops.launch_count()
        ops.channel_sum_multi(jobs)
        assert ops.launch_count() == n0 + 1
        for (x, out), ref in zip(jobs, refs):
            assert (out.double() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), (x.shape, (out.double() - ref).abs().max().item())
