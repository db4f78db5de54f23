"""GPU parity of the module-level API (models, losses, attack loops, training) against fixtures generated from the
UNMODIFIED reference (tests/golden/*.npz, made by tests/golden/make_golden.py) and against the CPU oracle."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

import synth
from oracle import spaa_oracle as O

pytestmark = pytest.mark.gpu

CAM_HW, PRJ_HW = (24, 32), (32, 32)
LABELS = {i: f"class{i}" for i in range(1000)}
SETUP = {"classifier_crop_sz": (24, 24), "prj_brightness": 0.5, "prj_im_sz": (32, 32)}


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, atol, rtol=0.0, what=""):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = ((a - b).abs() - rtol * b.abs())
    assert err.max().item() <= atol, f"{what}: max abs err {(a - b).abs().max().item():.3e} (atol {atol}, rtol {rtol})"


def close_per_sample(a, b, atol, what="", max_forks=1):
    """Free-running attack trajectories can fork on a threshold decision; samples are independent, so
    all but `max_forks` samples must match tightly."""
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().flatten(1).max(1)[0]
    bad = int((err > atol).sum())
    assert bad <= max_forks, f"{what}: {bad} samples differ (per-sample max err {err.tolist()})"


def make_pcnet(P, cam_hw, use_rough=True):
    from spaa_b200 import models
    wn = models.WarpingNet(out_size=tuple(cam_hw))
    sn = models.ShadingNetSPAA(use_rough=use_rough)
    m = models.PCNet(P["mask"], nn.DataParallel(wn), nn.DataParallel(sn), use_rough=use_rough)
    m.load_state_dict(P, strict=True)
    return m.to(dev())


def make_cpp(P, prj_hw):
    from spaa_b200 import models
    m = models.CompenNetPlusplus(nn.DataParallel(models.WarpingNet(out_size=tuple(prj_hw))), nn.DataParallel(models.CompenNet()))
    m.load_state_dict(P, strict=True)
    return m.to(dev())


class TinyClf:
    """Reference-convention classifier object around synth.TinyClassifier (fields used: model, input_sz)."""

    def __init__(self, seed, input_sz=(20, 20), device=None):
        self.model = synth.TinyClassifier(seed).to(device or dev())
        self.input_sz = input_sz

    def __call__(self, im, crop_sz):
        from spaa_b200.classifier import preprocess
        raw = self.model(preprocess(im, crop_sz, self.input_sz))
        p = torch.softmax(raw, 1).detach().cpu()
        ps, idx = p.sort(descending=True)
        return raw, ps.numpy(), idx.numpy()


def check_param_grads(golden, prefix, model, tol=2e-5, loose=None):
    """loose = (predicate on the parameter name, tolerance): parameters behind a ReLU mask that is known to sit on its threshold in this fixture."""
    seen = 0
    tol0 = tol
    for n, p in model.named_parameters():
        tol = loose[1] if (loose is not None and loose[0](n)) else tol0
        g = p.grad
        k = f"{prefix}_g_{n}"
        if k in golden:
            ref = T(golden[k])
            close(g, ref, tol * max(1.0, ref.abs().max().item()), 1e-4, k)
            seen += 1
        elif f"{prefix}_gsum_{n}" in golden:
            ref = float(golden[f"{prefix}_gsum_{n}"][0])
            scale = float(golden[f"{prefix}_gabs_{n}"][0]) if f"{prefix}_gabs_{n}" in golden else max(1.0, abs(ref))
            stol = 1e-5 if tol <= 2e-5 else 10 * tol      # (bf16x3: one ReLU flip on a toy-size image moves a SUM of gradient entries by up to 4e-4 relative, measured)
            assert abs(g.double().sum().item() - ref) <= stol * scale + 1e-6, (n, g.double().sum().item(), ref)
            if f"{prefix}_gabs_{n}" in golden:
                assert abs(g.double().abs().sum().item() - scale) <= stol * scale + 1e-6, n
            seen += 1
    assert seen > 10


# ---------------------------------------------------------------------------------------------------------

def test_warping_net_vs_reference(golden):
    from spaa_b200 import models, pytorch_tps
    g = golden("warp")
    P = synth.warping_params(21)
    sd = {k[len("warping_net."):]: v for k, v in P.items()}
    close(pytorch_tps.uniform_grid((6, 6)), g["uniform_grid"], 0)
    tg = pytorch_tps.tps_grid(P["warping_net.theta"].to(dev()), P["warping_net.ctrl_pts"].to(dev()), (1, 3, 12, 16))
    close(tg, g["tps_grid"], 2e-6, 0, "tps_grid")
    wn = models.WarpingNet(out_size=(12, 16))
    wn.load_state_dict(sd, strict=True)
    wn = wn.to(dev())
    x = synth.textured(22, "warp.x", (2, 3, 20, 20)).to(dev()).requires_grad_(True)
    y = wn(x)
    close(y, g["y"], 1e-5, 0, "warp y")
    cot = synth.randn(23, "warp.cot", y.shape).to(dev())
    (y * cot).sum().backward()
    close(x.grad, g["gx"], 1e-4, 1e-5, "gx")      # cotangents are O(4); fp32 noise of the sampling grid (1e-6) x 20 px
    for n, p in wn.named_parameters():
        ref = T(g["g_" + n])
        close(p.grad, ref, 3e-4 * max(1.0, ref.abs().max().item()), 1e-3, "g_" + n)
    with torch.no_grad():
        wn.simplify(x)
        close(wn.fine_grid, g["fine_grid"], 1e-5, 0, "fine_grid")
        close(wn(x), g["y"], 1e-5, 0, "simplified y")
        wn3 = models.WarpingNet(out_size=(12, 16), with_refine=False)
        wn3.load_state_dict({k: v for k, v in sd.items() if "grid_refine" not in k}, strict=True)
        close(wn3.to(dev())(x), g["y_norefine"], 1e-5, 0, "y_norefine")


def test_pcnet_and_compennetpp_vs_reference(golden):
    g = golden("models")
    P = synth.pcnet_params(31, CAM_HW)
    m = make_pcnet(P, CAM_HW)
    assert list(m.state_dict().keys()) == list(P.keys()) or set(m.state_dict().keys()) == set(P.keys())
    prj = synth.textured(32, "pc.prj", (2, 3, *PRJ_HW)).to(dev()).requires_grad_(True)
    scene = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).expand(2, -1, -1, -1).to(dev())
    y = m(prj, scene)
    close(y, g["pcnet_y"], 1e-5, 0, "pcnet y")
    cot = synth.randn(34, "pc.cot", y.shape).to(dev())
    (y * cot).sum().backward()
    close(prj.grad, g["pcnet_gprj"], 2e-5, 1e-4, "pcnet gprj")
    check_param_grads(g, "pcnet", m)
    with torch.no_grad():
        x = synth.textured(35, "sn.x", (2, 3, *CAM_HW)).to(dev())
        close(m.shading_net(x, scene, x * scene), g["shading_y"], 1e-5, 0, "shading y")
        Pn = synth.pcnet_params(36, CAM_HW, use_rough=False)
        close(make_pcnet(Pn, CAM_HW, use_rough=False)(prj, scene), g["pcnet_norough_y"], 1e-5, 0, "no-rough y")
        ms = make_pcnet(P, CAM_HW)
        ms.warping_net.simplify(prj)
        close(ms(prj, scene), g["pcnet_simplified_warp_y"], 1e-5, 0, "simplified-warp y")
    C = synth.compennet_pp_params(37)
    cm = make_cpp(C, PRJ_HW)
    cam = synth.textured(38, "cpp.cam", (2, 3, *CAM_HW)).to(dev()).requires_grad_(True)
    yc = cm(cam, scene)
    close(yc, g["cpp_y"], 1e-5, 0, "cpp y")
    cotc = synth.randn(39, "cpp.cot", yc.shape).to(dev())
    (yc * cotc).sum().backward()
    close(cam.grad, g["cpp_gcam"], 2e-5, 1e-4, "cpp gcam")
    check_param_grads(g, "cpp", cm)


def test_compute_loss_and_ssim_vs_reference(golden):
    from spaa_b200 import pytorch_ssim, train_network
    g = golden("loss")
    a0 = synth.textured(41, "loss.a", (2, 3, 24, 32))
    b = (a0 + synth.randn(42, "loss.b", a0.shape, 0.1)).clamp(0, 1).to(dev())
    for opt in ("l1", "l1+ssim", "l1+l2+ssim", "l2+huber", "ssim"):
        a = a0.clone().to(dev()).requires_grad_(True)
        loss, l2 = train_network.compute_loss(a, b, opt)
        loss.backward()
        key = opt.replace("+", "_")
        close(loss, g[key + "_loss"], 2e-6, 1e-5, key + " loss")
        close(l2, g[key + "_l2"], 1e-7, 1e-5, key + " l2")
        close(a.grad, g[key + "_g"], 3e-7, 2e-3, key + " grad")
    with torch.no_grad():
        a = a0.to(dev())
        close(pytorch_ssim.ssim(a, b), g["ssim_fn"], 2e-5, 0, "ssim()")
        close(pytorch_ssim.SSIM(size_average=False).to(dev())(a, b), g["ssim_per_sample"], 2e-5, 0, "SSIM(size_average=False)")
        close(pytorch_ssim.create_window(11, 1), g["window"], 1e-8, 0, "window")
    with pytest.raises(TypeError):
        train_network.compute_loss(a, b, "")


# ---------------------------------------------------------------------------------------------------------
# SPAA
# ---------------------------------------------------------------------------------------------------------

def _spaa_setup(seed_p=61, seed_s=62):
    P = synth.pcnet_params(seed_p, CAM_HW)
    m = make_pcnet(P, CAM_HW).eval()
    for p in m.parameters():
        p.requires_grad = False
    scene = synth.textured(seed_s, "spaa.scene", (1, 3, *CAM_HW))
    return P, m, scene


def test_spaa_teacher_forced_vs_oracle(golden):
    """Every iteration starts from the ORACLE's projector image; compares PCNet output, losses, masks and the update."""
    from spaa_b200 import projector_based_attack as pba
    P, m, scene = _spaa_setup()
    targets = [int(v) for v in golden("spaa")["targets"]]
    tiny_cpu = synth.TinyClassifier(1)
    iters = 10
    for loss_name, d_thr in (("camdE_caml2", 2.0), ("prjl2_caml2_camdE", 1.0)):
        otrace = []
        O.spaa_attack(lambda x, s: O.pcnet(P, x, s, CAM_HW), lambda im: O.classify(tiny_cpu, im, (24, 24), (20, 20)), targets, True, scene,
                      d_thr, loss_name, prj_hw=PRJ_HW, iters=iters, trace=otrace)
        forced = [t["prj_in"].to(dev()) for t in otrace]
        trace = []
        pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, d_thr, loss_name, dev(), SETUP, iters=iters, trace=trace, forced_prj=forced)
        n_col = n_flip_samples = n_bad = n_tot = 0
        for i, (a, o) in enumerate(zip(trace, otrace)):
            close(a["cam"], o["cam"], 1e-5, 0, f"it{i} cam")
            close(a["logits"], o["logits"], 2e-4, 1e-5, f"it{i} logits")
            hw = CAM_HW[0] * CAM_HW[1]
            close(a["stats"][:, 0] / hw, o["camde"], 2e-5, 1e-5, f"it{i} camdE")
            close(a["stats"][:, 1] / hw, o["caml2"], 1e-6, 1e-5, f"it{i} caml2")
            # decisions must agree except exactly at a threshold (p_top1 within 1e-4 of 0.9)
            p1 = torch.softmax(o["logits"], 1).max(1)[0]
            edge = (p1 - 0.9).abs() < 1e-4
            assert torch.equal(a["use_col"].cpu()[~edge], o["use_col"][~edge]), f"it{i} use_col"
            assert torch.equal(a["succ"].cpu(), o["succ"]), f"it{i} succ"
            n_col += int(o["use_col"].sum())
            same = (a["use_col"].cpu() == o["use_col"])
            a = {k: (v[same.to(v.device)] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == same.numel() else v) for k, v in a.items()}
            o = {k: (v[same] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == same.numel() else v) for k, v in o.items()}
            # the applied step: unit gradient of each sample's selected loss
            step_ref = o["prj_out"] - o["prj_in"]
            step_got = (a["prj_out"] - a["prj_in"]).cpu()
            # Typical agreement is 5e-7.  Rarely one sample shows a localised difference of a few 1e-4: a ReLU / clamp mask
            # of a pre-activation within rounding of 0 flips between the two fp32 evaluation orders and switches a
            # receptive field's worth of gradient on or off (tools/diag_spaa_step.py prints these events).  Allow at
            # most one such sample per iteration, bounded in size.
            err = (step_got - step_ref).abs()
            bad = (err > 2e-5 + 1e-3 * step_ref.abs()).flatten(1)
            n_flip_samples += int(bad.any(1).sum())
            n_bad += int(bad.sum()); n_tot += bad.numel()
            # (iteration 0 starts every sample from the same grey image: one such flip then shows in several samples at once)
            assert int(bad.any(1).sum()) <= (3 if i == 0 else 1) and err.max().item() <= 2e-3, f"it{i} step: {bad.sum(1).tolist()} max {err.max().item():.2e}"
            clean = ~bad.any(1)
            close(a["best_prj"], o["best_prj"], 2e-3, 0, f"it{i} best_prj")
            close(a["best_cam"], o["best_cam"], 1e-5, 0, f"it{i} best_cam")
        assert n_col > 0, "the stealth-loss branch was never exercised"
        assert n_flip_samples <= 5 and n_bad <= 0.01 * n_tot, (n_flip_samples, n_bad, n_tot)


def test_spaa_module_autograd_path_matches_fused_path(golden):
    """spaa() given an arbitrary callable (autograd through the nn.Module API) equals the fused engine path."""
    from spaa_b200 import projector_based_attack as pba
    P, m, scene = _spaa_setup()
    targets = [int(v) for v in golden("spaa")["targets"]]
    t1, t2 = [], []
    pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=8, trace=t1)
    forced = [t["prj_in"] for t in t1]
    pba.spaa(lambda x, s: m(x, s), TinyClf(1), LABELS, targets, True, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=8, trace=t2, forced_prj=forced)
    for i, (a, b) in enumerate(zip(t1, t2)):
        close(a["cam"], b["cam"], 1e-6, 0, f"it{i} cam")
        close(a["prj_out"], b["prj_out"], 1e-5, 0, f"it{i} prj_out")


def test_spaa_free_running_vs_reference(golden):
    from spaa_b200 import projector_based_attack as pba
    g = golden("spaa")
    P, m, scene = _spaa_setup()
    targets = [int(v) for v in g["targets"]]
    for tag, iters, loss, d_thr in (("t12", 12, "camdE_caml2", 2.0), ("t50", 50, "camdE", 3.0)):
        trace = []
        cam_best, prj_best = pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, d_thr, loss, dev(), SETUP, iters=iters, trace=trace)
        # Free-running trajectories fork on threshold decisions and then drift apart chaotically (the CPU oracle itself
        # forks on one of these samples, tests/test_oracle_golden.py); tight per-iteration parity is asserted by the
        # teacher-forced test above.  Here: most samples still track the reference, and the attack OUTCOME agrees.
        if iters <= 12:
            tol, forks = 2e-3, 2
            close_per_sample(trace[-1]["cam"], g[f"{tag}_cam_last"], tol, tag + " cam_last", forks)
            close_per_sample(torch.clamp(trace[-1]["prj_in"], 0, 1), g[f"{tag}_prj_last"], tol, tag + " prj_last", forks)
            close_per_sample(cam_best, g[f"{tag}_cam_best"], tol, tag + " cam_best", forks)
            close_per_sample(prj_best, g[f"{tag}_prj_best"], tol, tag + " prj_best", forks)
        # 50 free-running iterations: per-sample trajectories have decorrelated (and differ run to run: the grid_sample
        # scatter uses fp32 atomics); only the outcome is compared
        ref_l2 = torch.norm(T(g[f"{tag}_cam_best"]) - scene, dim=1).mean((1, 2))
        got_l2 = torch.norm(cam_best.cpu() - scene, dim=1).mean((1, 2))
        assert ((got_l2 - ref_l2).abs() <= 0.05 * ref_l2 + 1e-4).all(), (got_l2, ref_l2)
        # attack outcome: identical classifier top-1 on every attacked image (BASELINE.json north_star)
        clf = TinyClf(1)
        top_got = clf(cam_best, (24, 24))[2][:, 0]
        top_ref = clf(T(g[f"{tag}_cam_best"]).to(dev()), (24, 24))[2][:, 0]
        assert (top_got != top_ref).sum() <= 1, (top_got, top_ref)
    true_idx = int(g["u10_true_idx"])
    cam_best, prj_best = pba.spaa(m, TinyClf(1), LABELS, [true_idx], False, scene, 1.0, "prjl2_caml2_camdE", dev(), SETUP, iters=10)
    close(cam_best, g["u10_cam_best"], 2e-4, 0, "u10 cam_best")
    close(prj_best, g["u10_prj_best"], 2e-4, 0, "u10 prj_best")


# ---------------------------------------------------------------------------------------------------------
# PerC-AL + CompenNet++
# ---------------------------------------------------------------------------------------------------------

def test_percal_vs_reference(golden):
    from spaa_b200 import perc_al, projector_based_attack as pba
    g = golden("percal")
    scene = synth.textured(71, "pa.scene", (1, 3, *CAM_HW)).to(dev())
    clf = TinyClf(2)
    targets = [int(v) for v in g["targets"]]
    atk = perc_al.PerC_AL(device=dev(), max_iterations=15, alpha_l_init=1, alpha_c_init=0.5, confidence=0)
    xb = atk.adversary_projector(clf, scene.expand(8, -1, -1, -1), torch.tensor(targets), LABELS, 2.0, True, (24, 24))
    ref = T(g["t15_best"])
    # outputs are quantised to k/255: allow isolated one-level flips from rounding at .5 boundaries
    diff = (xb.cpu() - ref).abs()
    assert diff.max().item() <= 1.01 / 255 and (diff > 1e-6).float().mean().item() < 2e-3, (diff.max().item(), (diff > 1e-6).float().mean().item())
    true_idx = int(g["u15_true_idx"])
    atk = perc_al.PerC_AL(device=dev(), max_iterations=15, alpha_l_init=1, alpha_c_init=0.5, confidence=40)
    xu = atk.adversary_projector(clf, scene, torch.tensor([true_idx]), LABELS, 2.0, False, (24, 24))
    diff = (xu.cpu() - T(g["u15_best"])).abs()
    assert diff.max().item() <= 1.01 / 255 and (diff > 1e-6).float().mean().item() < 2e-3
    C = synth.compennet_pp_params(72)
    cm = make_cpp(C, PRJ_HW).eval()
    cam_best, prj_best = pba.perc_al_compennet_pp(cm, clf, LABELS, [true_idx], False, scene, 2.0, dev(), SETUP)
    diff = (cam_best.cpu() - T(g["full_cam_best"])).abs()
    assert diff.max().item() <= 1.01 / 255 and (diff > 1e-6).float().mean().item() < 2e-3
    close(prj_best, g["full_prj_best"], 2e-2, 0, "perc-al prj_best")     # CompenNet++ of a quantised image with rare 1/255 flips
    with pytest.raises(ValueError):
        atk.adversary_projector(clf, scene + 1.0, torch.tensor([true_idx]), LABELS, 2.0, False, (24, 24))


def test_percal_teacher_free_first_iterations_vs_oracle(golden):
    from spaa_b200 import perc_al
    scene = synth.textured(71, "pa.scene", (1, 3, *CAM_HW))
    tiny_cpu = synth.TinyClassifier(2)
    targets = torch.tensor([int(v) for v in golden("percal")["targets"]])
    otrace, trace = [], []
    O.perc_al_attack(lambda im: O.classify(tiny_cpu, im, (24, 24), (20, 20)), scene.expand(8, -1, -1, -1), targets, 2.0, True,
                     max_iterations=6, trace=otrace)
    atk = perc_al.PerC_AL(device=dev(), max_iterations=6, alpha_l_init=1, alpha_c_init=0.5, confidence=0)
    atk.adversary_projector(TinyClf(2), scene.expand(8, -1, -1, -1).to(dev()), targets, LABELS, 2.0, True, (24, 24), trace=trace)
    for i, (a, o) in enumerate(zip(trace, otrace)):
        close(a["delta"], o["delta"], 5e-5, 0, f"it{i} delta")
        close(a["dis"], o["dis"], 1e-2, 1e-4, f"it{i} dis")
        assert torch.equal(a["use_col"].cpu(), o["use_col"]) and torch.equal(a["isadv"].cpu(), o["isadv"]), f"it{i} masks"


# ---------------------------------------------------------------------------------------------------------
# training
# ---------------------------------------------------------------------------------------------------------

def _check_after(golden, prefix, model, P0, tol):
    """Parameters after a few Adam steps.  Adam's first steps move every element by ~lr * sign(g): elements whose
    gradient is ~0 get a noise-determined sign, so a few elements may legitimately differ by up to 2*lr; require 99% of
    each tensor within `tol` and every element within 2.5 * lr_max * n_steps."""
    seen = 0
    for k, v in model.state_dict().items():
        if f"{prefix}_after_{k}" in golden:
            ref = T(golden[f"{prefix}_after_{k}"]).double()
            err = (v.detach().cpu().double() - ref).abs().flatten()
            if err.numel():
                q = torch.quantile(err, 0.99).item() if err.numel() > 100 else err.median().item()
                assert q <= tol and err.max().item() <= 0.08, (k, q, err.max().item())
            seen += 1
        elif f"{prefix}_afterdelta_{k}" in golden:
            ref = float(golden[f"{prefix}_afterdelta_{k}"][0])
            got = (v.cpu() - P0[k]).double().abs().sum().item()
            assert abs(got - ref) <= 5e-3 * ref + 1e-6, (k, got, ref)
            seen += 1
    assert seen > 20


def test_train_pcnet_and_compennetpp_vs_reference(golden):
    from spaa_b200 import train_network as tn
    g = golden("train")
    N = 6
    P = synth.pcnet_params(81, CAM_HW)
    m = nn.DataParallel(make_pcnet(P, CAM_HW), device_ids=[0])
    prj_train = synth.textured(82, "tr.prj", (N, 3, *PRJ_HW))
    scene = synth.textured(83, "tr.scene", (1, 3, *CAM_HW))
    cam_train = synth.textured(84, "tr.cam", (N, 3, *CAM_HW))
    cfg = tn.AttrDict(device="cuda:0", data_root=None, setup_name="synth", model_name="PCNet", num_train=N, batch_size=4, max_iters=3, lr=1e-3,
                      lr_drop_ratio=0.2, lr_drop_rate=800, l2_reg=1e-4, plot_on=False, train_plot_rate=50, valid_rate=200, loss="l1+ssim")
    random.seed(5)
    tn.train_pcnet(m, dict(cam_scene=scene, cam_train=cam_train, prj_train=prj_train, mask=P["mask"]), None, cfg, verbose=False)
    close(cfg["loss_history"][:, 0], g["pcnet_losses"], 1e-4, 0, "pcnet losses")      # the reference prints 4 decimals
    _check_after(g, "pcnet", m.module, P, 3e-4)      # Adam divides by sqrt(v): near-zero gradients amplify fp32 noise
    C = synth.compennet_pp_params(85)
    cm = nn.DataParallel(make_cpp(C, PRJ_HW), device_ids=[0])
    cfg2 = tn.AttrDict(dict(cfg))
    cfg2.model_name, cfg2.loss = "CompenNet++", "l1+ssim"
    random.seed(6)
    tn.train_compennet_pp(cm, dict(cam_scene=scene, cam_train=cam_train, prj_train=prj_train), None, cfg2, verbose=False)
    close(cfg2["loss_history"][:, 0], g["cpp_losses"], 1e-4, 0, "cpp losses")
    _check_after(g, "cpp", cm.module, C, 3e-4)


def test_evaluate_model_and_metrics():
    from spaa_b200 import train_network as tn, utils as ut
    P = synth.pcnet_params(81, CAM_HW)
    m = nn.DataParallel(make_pcnet(P, CAM_HW), device_ids=[0])
    N = 5
    prj = synth.textured(91, "ev.prj", (N, 3, *PRJ_HW))
    cam = synth.textured(92, "ev.cam", (N, 3, *CAM_HW))
    scene = synth.textured(93, "ev.scene", (1, 3, *CAM_HW)).expand(N, -1, -1, -1)
    psnr, rmse, ssim, infer = tn.evaluate_model(m, dict(cam_scene=scene, cam_valid=cam, prj_valid=prj), chunk_sz=2)
    ref = O.pcnet(P, prj, scene, CAM_HW)
    close(infer, ref, 1e-5, 0, "evaluate_model inference")
    mse = ((ref - cam) ** 2).mean().item()
    assert abs(rmse - (mse * 3) ** 0.5) < 2e-2 and abs(ssim - O.ssim_index(ref, cam).item()) < 1e-3
    d = ut.calc_img_dists(ref.to(dev()), cam.to(dev()))
    assert abs(d[5] - O.mean_delta_e(ref, cam)) < 1e-3 and abs(d[3] - torch.norm(ref - cam, dim=1).mean().item() * 255) < 1e-2


# ---------------------------------------------------------------------------------------------------------
# bf16 tensor-core path (BASELINE.json: within 2e-3 max-abs of the fp32 result, identical classifier top-1)
# ---------------------------------------------------------------------------------------------------------

# Measured on B200 (this fixture): PCNet output max-abs error 2.9e-3 with pure bf16 storage, 3.4e-4 with fp16 forward
# activations -- so 'fp16' is the mode that meets BASELINE.json's 2e-3 bar and the one bench.py reports; 'bf16' is kept
# as the wider-range variant and bounded at 4e-3.
MODES = [("bf16", 4e-3, 0.12, 0.95), ("fp16", 2e-3, 0.04, 0.98)]      # precision, output tol, gradient rel-F tol, min update cosine


@pytest.mark.parametrize("precision,tol,gtol,mincos", MODES)
def test_pcnet_16bit_tensor_core_path_vs_reference(golden, precision, tol, gtol, mincos):
    """Tensor-core path vs the fp32 reference fixture (tests/golden/models.npz)."""
    from spaa_b200 import models, ops
    g = golden("models")
    P = synth.pcnet_params(31, CAM_HW)
    m = models.set_precision(make_pcnet(P, CAM_HW), precision)
    prj = synth.textured(32, "pc.prj", (2, 3, *PRJ_HW)).to(dev()).requires_grad_(True)
    scene = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).expand(2, -1, -1, -1).to(dev())
    probe = ops.set_probe(lambda kind, spec: kind.endswith("_tc"))
    y = m(prj, scene)
    cot = synth.randn(34, "pc.cot", y.shape).to(dev())
    (y * cot).sum().backward()
    n_tc = len(probe["events"])
    ops.set_probe(None)
    assert n_tc >= 20, f"only {n_tc} launches went through the tcgen05 kernel"
    err = (y.detach().cpu() - T(g["pcnet_y"])).abs()
    print(f"16bit[{precision}] PCNet output: max abs err {err.max().item():.2e}, mean {err.mean().item():.2e}")
    assert err.max().item() <= tol, err.max().item()
    gref = T(g["pcnet_gprj"]).double()
    got = prj.grad.cpu().double()
    rel = ((got - gref).norm() / gref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(1), gref.flatten(1), dim=1)
    print(f"16bit[{precision}] d/dprj: relative Frobenius error {rel:.3e}, per-sample cosine {cos.tolist()}")
    assert rel <= gtol and (cos >= mincos).all(), (rel, cos.tolist())


@pytest.mark.parametrize("precision,tol,gtol,mincos", MODES)
def test_spaa_16bit_teacher_forced_outcome(golden, precision, tol, gtol, mincos):
    """Tensor-core PCNet inside the attack loop: per-iteration camera image close to the fp32 oracle, identical top-1."""
    from spaa_b200 import models, projector_based_attack as pba
    P, m, scene = _spaa_setup()
    targets = [int(v) for v in golden("spaa")["targets"]]
    tiny_cpu = synth.TinyClassifier(1)
    otrace, trace = [], []
    O.spaa_attack(lambda x, s: O.pcnet(P, x, s, CAM_HW), lambda im: O.classify(tiny_cpu, im, (24, 24), (20, 20)), targets, True, scene,
                  2.0, "camdE_caml2", prj_hw=PRJ_HW, iters=8, trace=otrace)
    pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=8, trace=trace,
             forced_prj=[t["prj_in"].to(dev()) for t in otrace], precision=precision)
    models.set_precision(m, "fp32")
    worst = 1.0
    for i, (a, o) in enumerate(zip(trace, otrace)):
        close(a["cam"], o["cam"], tol, 0, f"it{i} cam ({precision})")
        assert torch.equal(a["logits"].argmax(1).cpu(), o["logits"].argmax(1)), f"it{i} top-1"
        step_ref = o["prj_out"] - o["prj_in"]
        same = (a["use_col"].cpu() == o["use_col"])
        cos = torch.nn.functional.cosine_similarity((a["prj_out"] - a["prj_in"]).cpu().flatten(1), step_ref.flatten(1), dim=1)
        worst = min(worst, cos[same].min().item())
        assert (cos[same] > mincos).all(), f"it{i} update direction cos {cos.tolist()}"
    print(f"16bit[{precision}] teacher-forced: worst update-direction cosine {worst:.4f}")


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_spaa_16bit_free_running_success_rate(precision):
    """BASELINE.json: the 16-bit path must agree with fp32 on attack success within +-1 image per 100 (free-running, so
    individual trajectories may fork at a threshold; the COUNT is the invariant) and classify every attacked image
    (the returned cam_infer_best) to the same top-1 as the fp32 classifier does on the same image."""
    from spaa_b200 import models, projector_based_attack as pba
    P, m, scene = _spaa_setup()
    targets = [(37 * i + 11) % 1000 for i in range(100)]
    tiny_cpu, clf = synth.TinyClassifier(1), TinyClf(1)
    iters, d_thr = 40, 2.0
    ocam, oprj = O.spaa_attack(lambda x, s: O.pcnet(P, x, s, CAM_HW), lambda im: O.classify(tiny_cpu, im, (24, 24), (20, 20)), targets, True,
                               scene, d_thr, "camdE_caml2", prj_hw=PRJ_HW, iters=iters)
    cam, prj = pba.spaa(m, clf, LABELS, targets, True, scene, d_thr, "camdE_caml2", dev(), SETUP, iters=iters, precision=precision)
    models.set_precision(m, "fp32")
    tgt = torch.tensor(targets)

    def success(cam_best):
        lg = O.classify(tiny_cpu, cam_best.cpu(), (24, 24), (20, 20))[0].detach()
        return lg.argmax(1) == tgt, lg
    s_ref, _ = success(ocam)
    s_got, lg_cpu = success(cam)
    print(f"16bit[{precision}] free-running: fp32 oracle {int(s_ref.sum())}/100 successful attacks, tensor-core path {int(s_got.sum())}/100")
    assert 0 < int(s_ref.sum()) < 100, "fixture must produce a mix of successes and failures"
    assert abs(int(s_got.sum()) - int(s_ref.sum())) <= 1
    # top-1 of every attacked image: device classifier on our output == fp32 CPU classifier on the same image
    from spaa_b200.classifier import device_logits
    lg_dev = device_logits(clf, cam, (24, 24))
    assert torch.equal(lg_dev.argmax(1).cpu(), lg_cpu.argmax(1))


def test_train_pcnet_bf16_tensor_core_tracks_fp32():
    """Mixed-precision training (bf16 activations / gradients on tcgen05 incl. the backward-weight kernel, fp32 master weights):
    the loss trajectory and the parameter update after a few steps stay close to the exact fp32 mode on the same batches
    (this also guards the packed-weight cache: a stale cache makes the 16-bit run ignore the optimizer)."""
    from spaa_b200 import models, train_network as tn
    N, hw, phw = 6, (48, 64), (64, 64)
    P = synth.pcnet_params(81, hw)
    prj_train = synth.textured(82, "trb.prj", (N, 3, *phw))
    scene = synth.textured(83, "trb.scene", (1, 3, *hw))
    cam_train = synth.textured(84, "trb.cam", (N, 3, *hw))
    from spaa_b200 import ops
    res, n_wg = {}, {}
    for prec in ("fp32", "bf16", "fp16"):
        probe = ops.set_probe(lambda kind, spec: kind == "bwd_weight_tc")
        wn = models.WarpingNet(out_size=hw)
        m = models.PCNet(P["mask"], nn.DataParallel(wn), nn.DataParallel(models.ShadingNetSPAA()))
        m.load_state_dict(P, strict=True)
        m = nn.DataParallel(models.set_precision(m.to(dev()), prec), device_ids=[0])
        cfg = tn.AttrDict(device="cuda:0", data_root=None, model_name="PCNet", num_train=N, batch_size=4, max_iters=6, lr=1e-3, lr_drop_ratio=0.2,
                          lr_drop_rate=800, l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, iter_offset=401, save_checkpoint=False)
        random.seed(5)
        tn.train_pcnet(m, dict(cam_scene=scene, cam_train=cam_train, prj_train=prj_train, mask=P["mask"]), None, cfg, verbose=False)
        n_wg[prec] = len(probe["events"])          # (eager steps only: launches replayed from the CUDA graph are not host-timed)
        ops.set_probe(None)
        res[prec] = (cfg["loss_history"][:, 0].cpu(), {k: v.detach().cpu().double() for k, v in m.module.state_dict().items()})
    l32 = res["fp32"][0]
    assert (l32[0] - l32[-1]).item() > 1e-3, "fixture must make progress"
    # both 16-bit modes run EVERY backward-weight on the tcgen05 kernel ('fp16': fp16 forward activations re-rounded to bf16 for that launch): 17 conv
    # layers per eager step
    assert n_wg["fp32"] == 0 and n_wg["bf16"] >= 3 * 14 and n_wg["fp16"] >= 3 * 14, n_wg
    for prec in ("bf16", "fp16"):
        l16 = res[prec][0]
        print("losses fp32", l32.tolist(), prec, l16.tolist())
        close(l16, l32, 5e-3 if prec == "bf16" else 2e-3, 0, prec + " loss trajectory")
        assert abs((l16[0] - l16[-1]).item() - (l32[0] - l32[-1]).item()) < 0.25 * (l32[0] - l32[-1]).item(), prec + " run does not follow the optimizer"
        for k in ("shading_net.conv4.weight", "shading_net.conv1_s.weight", "shading_net.conv6.weight", "shading_net.transConv1.weight"):
            d32 = (res["fp32"][1][k] - P[k].double()).flatten()
            d16 = (res[prec][1][k] - P[k].double()).flatten()
            cos = torch.nn.functional.cosine_similarity(d32, d16, dim=0).item()
            assert cos > 0.9, (prec, k, cos)


@pytest.mark.parametrize("prec,ltol", [("fp32", 5e-6), ("bf16", 2e-3)])
def test_train_pcnet_cuda_graph_replay_matches_eager_steps(prec, ltol):
    """The training step replayed from its CUDA graph (steps 4.. of a run: batch gather from the staged indices, forward, fused loss,
    backward, Adam with the device-resident scheduler row) follows the trajectory of launching every step eagerly, across the
    L1 -> L1+SSIM phase switch (a second graph).  Run-to-run the eager loop itself differs by ~1e-6 (atomic accumulation order) and Adam
    amplifies that step by step (1e-2 learning rate on the affine parameters, six images), so only the first steps after the first
    capture are held tightly; the rest of the trajectory must stay close and make the same progress."""
    from spaa_b200 import models, train_network as tn
    N, hw, phw = 6, (48, 64), (64, 64)
    P = synth.pcnet_params(81, hw)
    prj_train = synth.textured(82, "trg.prj", (N, 3, *phw))
    scene = synth.textured(83, "trg.scene", (1, 3, *hw))
    cam_train = synth.textured(84, "trg.cam", (N, 3, *hw))
    res = {}
    for graph in (False, True):
        m = models.PCNet(P["mask"], nn.DataParallel(models.WarpingNet(out_size=hw)), nn.DataParallel(models.ShadingNetSPAA()))
        m.load_state_dict(P, strict=True)
        m = nn.DataParallel(models.set_precision(m.to(dev()), prec), device_ids=[0])
        cfg = tn.AttrDict(device="cuda:0", data_root=None, model_name="PCNet", num_train=N, batch_size=4, max_iters=12, lr=1e-3, lr_drop_ratio=0.2,
                          lr_drop_rate=800, l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, iter_offset=396, save_checkpoint=False, graph=graph)
        random.seed(5)
        tn.train_pcnet(m, dict(cam_scene=scene, cam_train=cam_train, prj_train=prj_train, mask=P["mask"]), None, cfg, verbose=False)
        res[graph] = cfg["loss_history"].cpu().double()
    le, lg = res[False], res[True]
    print("eager", le[:, 0].tolist(), "graph", lg[:, 0].tolist())
    assert torch.isfinite(lg).all()
    # (measured with tools/graph_noise_probe.py: eager-vs-eager and graph-vs-eager agree to 1e-7 up to step 4, then both drift apart
    # 5e-5, 8e-5, 6e-4 ... per step -- the same run-to-run noise, amplified by Adam)
    close(lg[:4], le[:4], ltol, ltol, f"{prec}: first steps (3 eager, then capture + first replay), graph vs eager")
    close(lg[:, 0], le[:, 0], 3e-2, 0, f"{prec}: loss trajectory, graph vs eager")
    # the L1+SSIM phase (iterations 5..11: three eager steps, the second capture, replays) keeps descending like the eager run
    assert abs((lg[5, 0] - lg[-1, 0]) - (le[5, 0] - le[-1, 0])).item() <= 0.5 * abs((le[5, 0] - le[-1, 0]).item()) + 5e-3


def test_flat_adam_graph_replay_matches_eager_updates():
    """FlatAdam.advance() + a captured apply() == FlatAdam.step() launched eagerly, bit for bit, over steps that cross a learning-rate
    milestone: the replayed launch must pick up each step's bias corrections and learning rates from the device-resident row."""
    from spaa_b200 import train_network as tn
    torch.manual_seed(3)
    shapes = [(7, 5), (11,), (3, 2, 2)]
    grads = [torch.randn(12, *sh, device=dev()) for sh in shapes]

    def run(graph: bool):
        ps = [torch.nn.Parameter(torch.linspace(-1, 1, int(np.prod(sh)), device=dev()).reshape(sh).clone()) for sh in shapes]
        opt = tn.FlatAdam([(ps[:2], 1e-2, 0.0, [4], 0.2), (ps[2:], 1e-3, 1e-4, [7], 0.5)], max_steps=12)
        g = None
        for t in range(12):
            opt.zero_grad()
            for p, gr in zip(ps, grads):
                p.grad.copy_(gr[t])
            if not graph:
                opt.step(grad_scale=0.5)
                continue
            opt.advance()
            if g is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    opt.apply(grad_scale=0.5)
            g.replay()
        torch.cuda.synchronize()
        return [p.detach().clone() for p in ps]

    for a, b in zip(run(False), run(True)):
        assert torch.equal(a, b)


def test_run_attack_sweep_matches_direct_calls_and_writes_the_reference_layout(tmp_path):
    """The sharded sweep driver (job loop of run_projector_based_attack, projector_based_attack.py:83-141): every job = targeted batch +
    untargeted attack on the scene's own top-1; results equal the direct spaa() calls, land under the reference's folder names as
    img_0001.. (targeted first, untargeted last), and the ranks' shards partition the jobs."""
    from spaa_b200 import projector_based_attack as pba, utils as ut
    from spaa_b200.classifier import device_logits
    P, m, scene = _spaa_setup()
    scene2 = synth.textured(63, "sweep.scene", (1, 3, *CAM_HW))
    clf = TinyClf(1)
    targets = [808, 969, 116]
    jobs = [dict(model=m, classifier=clf, classifier_name="tiny", cam_scene=sc, target_idx=targets, stealth_loss=sl, d_thr=thr, setup_info=SETUP,
                 setup_path=str(tmp_path / name)) for name, sc, sl, thr in (("s1", scene, "camdE_caml2", 2.0), ("s2", scene2, "caml2", 3.0), ("s1", scene, "camdE", 2.0))]
    done = {}
    for rank in range(2):
        for ji, r in pba.run_attack_sweep(jobs, dev(), iters=4, rank=rank, world=2):
            assert ji % 2 == rank and ji not in done
            done[ji] = r
    assert sorted(done) == [0, 1, 2]
    cfg_str = pba.to_attacker_cfg_str("SPAA")[0]
    for ji, job in enumerate(jobs):
        r = done[ji]
        sc = job["cam_scene"].to(dev())
        true_idx = int(device_logits(clf, sc, SETUP["classifier_crop_sz"]).argmax(1)[0])
        assert r["true_idx"] == true_idx and r["n_targets"] == 3 and r["cam_infer"].shape == (4, 3, *CAM_HW) and r["prj_adv"].shape == (4, 3, *PRJ_HW)
        cam_t, prj_t = pba.spaa(m, clf, LABELS, targets, True, sc, job["d_thr"], job["stealth_loss"], dev(), SETUP, iters=4)
        cam_u, prj_u = pba.spaa(m, clf, LABELS, [true_idx], False, sc, job["d_thr"], job["stealth_loss"], dev(), SETUP, iters=4)
        close(r["cam_infer"], torch.cat((cam_t, cam_u)).cpu(), 1e-5, 0, f"job {ji} cam")       # (scatter-add order: not bit-exact run to run)
        close(r["prj_adv"], torch.cat((prj_t, prj_u)).cpu(), 1e-4, 0, f"job {ji} prj")
        folder = os.path.join(cfg_str, job["stealth_loss"], str(job["d_thr"]), "tiny")
        for sub, ref in (("cam/infer/adv", r["cam_infer"]), ("prj/adv", r["prj_adv"])):
            d = os.path.join(job["setup_path"], sub, folder)
            assert sorted(os.listdir(d)) == [f"img_{k:04d}.png" for k in range(1, 5)]
            back = ut.torch_imread(os.path.join(d, "img_0004.png"))                             # the untargeted result is stored last (:137-138)
            assert (back - ref[3].clamp(0, 1)).abs().max().item() <= 1.0 / 255 + 1e-6


def test_spaa_deterministic_mode_matches_default():
    """deterministic=True swaps the scatter-add backward of the warp (atomics) for the gather through the per-attack CSR adjoint map with the
    fused norm: same attack within fp32 rounding, for both stealth-loss families (with / without the projector L2 term)."""
    from spaa_b200 import projector_based_attack as pba
    P, m, scene = _spaa_setup()
    targets = [808, 969, 116, 786]
    for loss_name in ("camdE_caml2", "prjl2_caml2_camdE"):
        pba.clear_engines()
        cam_a, prj_a = pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, loss_name, dev(), SETUP, iters=6)
        cam_b, prj_b = pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, loss_name, dev(), SETUP, iters=6, deterministic=True)
        close_per_sample(prj_b, prj_a, 2e-4, f"{loss_name}: projector images")
        close_per_sample(cam_b, cam_a, 2e-4, f"{loss_name}: camera images")


def test_engines_sharing_the_graph_memory_pool_do_not_disturb_each_other(monkeypatch):
    """Two cached engines (targeted / untargeted, as in every sweep job) capture into ONE CUDA-graph memory pool (ops.graph_pool): the second capture
    may take blocks the first one released, so their replays are interleaved here job by job -- results must equal the ones from private pools up to
    the run-to-run noise of the fp32 atomics that remain in deterministic mode (measured 7e-7 after 8 iterations; a clobbered buffer shows as 1e-2 or
    worse).  Also: clear_engines() releases the pool and the next capture starts a new one."""
    from spaa_b200 import ops, projector_based_attack as pba
    P, m, scene = _spaa_setup()
    clf = TinyClf(1)
    jobs = [([808, 969, 116, 786], True), ([5], False), ([1, 2, 3, 4], True), ([7], False)]

    def run():
        pba.clear_engines()
        out = []
        for targets, targeted in jobs:
            cam, prj = pba.spaa(m, clf, LABELS, targets, targeted, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=8, deterministic=True, graph=True)
            out.append((cam.clone(), prj.clone()))
        return out

    def same(a, b, what):
        for j, ((ca, pa), (cb, pb)) in enumerate(zip(a, b)):
            dc, dp = (ca - cb).abs().max().item(), (pa - pb).abs().max().item()
            assert dc <= 2e-5 and dp <= 2e-5, (what, j, dc, dp)

    monkeypatch.setenv("SPAA_GRAPH_POOL", "0")
    ref = run()
    same(ref, run(), "private pools, run to run")
    monkeypatch.setenv("SPAA_GRAPH_POOL", "1")
    got = run()
    assert len(ops._graph_pools) == 1
    same(ref, got, "shared pool")
    pba.clear_engines()
    assert len(ops._graph_pools) == 0
    same(ref, run(), "a fresh shared pool after the release")
    pba.clear_engines()


def test_attack_engine_cache_sees_training_updates_and_simplify():
    """spaa() -> train_pcnet() on the same model object -> spaa() must not reuse the cached engine (its captured CUDA graph holds the
    old packed weights, its grid / skip activations were computed from the old parameters); FlatAdam's raw kernel does not bump version counters."""
    from spaa_b200 import projector_based_attack as pba, train_network as tn, models
    P = synth.pcnet_params(61, CAM_HW)
    m = make_pcnet(P, CAM_HW)
    models.set_precision(m, "fp16")
    scene = synth.textured(62, "spaa.scene", (1, 3, *CAM_HW))
    targets = [3, 5, 7, 11]
    pba.clear_engines()

    def attack():
        m.eval()
        for p in m.parameters():
            p.requires_grad = False
        return pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=4)
    cam0, _ = attack()
    n_engines = len(pba._ENGINES)
    for p in m.parameters():
        p.requires_grad = True
    N = 6
    cfg = tn.AttrDict(device="cuda:0", data_root=None, setup_name="synth", model_name="PCNet", num_train=N, batch_size=4, max_iters=2, lr=1e-2,
                      lr_drop_ratio=0.2, lr_drop_rate=800, l2_reg=1e-4, plot_on=False, train_plot_rate=50, valid_rate=10 ** 9, loss="l1+ssim",
                      save_checkpoint=False)
    random.seed(5)
    tn.train_pcnet(nn.DataParallel(m, device_ids=[0]), dict(cam_scene=scene, cam_train=synth.textured(84, "tr.cam", (N, 3, *CAM_HW)),
                                                          prj_train=synth.textured(82, "tr.prj", (N, 3, *PRJ_HW)), mask=P["mask"]), None, cfg, verbose=False)
    cam1, _ = attack()                                   # same model object, updated weights
    pba.clear_engines()
    cam_fresh, _ = attack()
    close(cam1, cam_fresh, 1e-6, 0, "attack after training vs a fresh engine")
    assert (cam1 - cam0).abs().max().item() > 1e-4, "training did not change the model: the test does not exercise the cache"
    # simplify() adds buffers (version 0): a different key as well
    m.warping_net.simplify(torch.zeros(1, 3, *PRJ_HW, device=dev()))
    cam2, _ = attack()
    close(cam2, cam_fresh, 1e-5, 0, "attack with a simplified WarpingNet")
    pba.clear_engines()
    models.set_precision(m, "fp32")


# ---------------------------------------------------------------------------------------------------------
# API corners pinned to the unmodified reference (tests/golden/extra.npz, make_golden.py gen_extra)
# ---------------------------------------------------------------------------------------------------------

def test_tps_full_form_and_batched_vs_reference(golden):
    """pytorch_tps.tps / tps_grid (pytorch_tps.py:29-106): full-form theta (T+3 rows), a batch of two TPS, per-sample control points, gradient."""
    from spaa_b200 import pytorch_tps
    g = golden("extra")
    nT = 36
    ctrl = pytorch_tps.uniform_grid((6, 6)).view(-1, 2).to(dev())
    th_full = synth.randn(101, "tps.full", (2, nT + 3, 2), 0.01).to(dev()).requires_grad_(True)
    th_red = synth.randn(102, "tps.red", (2, nT + 2, 2), 0.01).to(dev())
    ctrl_b = torch.stack((ctrl.cpu(), (ctrl.cpu() + synth.randn(103, "tps.ctrl", ctrl.shape, 0.01)).clamp(0, 1))).to(dev())
    gf = pytorch_tps.tps_grid(th_full, ctrl, (2, 3, 12, 16))
    close(gf, g["tps_grid_full"], 2e-6, 0, "tps_grid full form")
    close(pytorch_tps.tps_grid(th_red, ctrl_b, (2, 3, 12, 16)), g["tps_grid_red_b"], 2e-6, 0, "tps_grid batched ctrl")
    g3 = torch.ones(2, 12, 16, 3, device=dev())
    g3[..., 1] = torch.linspace(0, 1, 16)
    g3[..., 2] = torch.linspace(0, 1, 12).unsqueeze(-1)
    close(pytorch_tps.tps(th_full, ctrl, g3), g["tps_z_full"], 2e-6, 0, "tps() full form")
    cot = synth.randn(104, "tps.cot", (2, 12, 16, 2)).to(dev())
    (gf * cot).sum().backward()
    ref = T(g["tps_grid_full_gtheta"])
    close(th_full.grad, ref, 2e-5 * max(1.0, ref.abs().max().item()), 1e-4, "d tps_grid / d theta (full form)")
    with pytest.raises(NotImplementedError):
        pytorch_tps.tps(th_full, ctrl, g3 * 0.9)


def test_ssim_mask_and_weights_branches_vs_reference(golden):
    """pytorch_ssim/__init__.py:53-67: pixel weights, boolean mask with size_average, float mask per sample -- values and d/d img1."""
    from spaa_b200 import pytorch_ssim
    g = golden("extra")
    a0 = synth.textured(111, "sx.a", (2, 3, 24, 32))
    b = (a0 + synth.randn(112, "sx.b", a0.shape, 0.1)).clamp(0, 1).to(dev())
    w, mb = T(g["ssim_w"]).to(dev()), T(g["ssim_mask"]).to(dev())
    for tag, kw, avg in (("w_avg", dict(weights=w), True), ("m_avg", dict(mask=mb), True), ("wm_avg", dict(mask=mb, weights=w), True),
                         ("m_per", dict(mask=mb.float()), False), ("w_per", dict(weights=w), False)):
        a = a0.clone().to(dev()).requires_grad_(True)
        v = pytorch_ssim.SSIM(size_average=avg).to(dev())(a, b, **kw)
        v.sum().backward()
        close(v, g["ssim_" + tag], 2e-5, 0, "ssim " + tag)
        ref = T(g["ssim_" + tag + "_g"])
        close(a.grad, ref, 2e-3 * ref.abs().max().item(), 2e-3, "ssim grad " + tag)


def test_simplify_paths_vs_reference(golden):
    """ShadingNetSPAA.simplify / PCNet.simplify / CompenNetPlusplus.simplify (models.py:149-161, 268-277, 330-333, 199-202)."""
    g = golden("extra")
    Pn = synth.pcnet_params(36, CAM_HW, use_rough=False)
    m = make_pcnet(Pn, CAM_HW, use_rough=False).eval()
    prj = synth.textured(32, "pc.prj", (2, 3, *PRJ_HW)).to(dev())
    scene1 = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).to(dev())
    with torch.no_grad():
        x = synth.textured(35, "sn.x", (2, 3, *CAM_HW)).to(dev())
        close(m.shading_net(x, scene1.expand(2, -1, -1, -1)), g["shading_norough_y"], 1e-5, 0, "shading (no rough) y")
        m.shading_net.simplify(scene1)
        close(m.shading_net.res1_s, g["res1_s"], 1e-5, 0, "res1_s")
        close(m.shading_net.res4_s, g["res4_s"], 1e-5, 0, "res4_s")
        assert "res1_s" in m.shading_net.state_dict()                        # buffers appear in the state dict after simplify, like the reference's
        close(m.shading_net(x, scene1.expand(2, -1, -1, -1)), g["shading_simplified_y"], 1e-5, 0, "simplified shading y")
        m2 = make_pcnet(Pn, CAM_HW, use_rough=False).eval()
        m2.simplify(scene1)
        close(m2.warping_net.fine_grid, g["pcnet_simplified_grid"], 1e-5, 0, "simplified fine_grid")
        close(m2(prj, scene1.expand(2, -1, -1, -1)), g["pcnet_simplified_y"], 1e-5, 0, "simplified PCNet y")
        C = synth.compennet_pp_params(37)
        cm = make_cpp(C, PRJ_HW).eval()
        cam = synth.textured(38, "cpp.cam", (2, 3, *CAM_HW)).to(dev())
        cm.simplify(scene1)
        close(cm(cam, scene1.expand(2, -1, -1, -1)), g["cpp_simplified_y"], 1e-5, 0, "simplified CompenNet++ y")


def test_percal_adversary_original_variant_vs_reference(golden):
    """PerC_AL.adversary (perc_al/__init__.py:53-131): the digital attack against a bare model fed (x - 0.5) / 0.5; targeted, untargeted and the
    confidence-40 untargeted variant.  Outputs are quantised to k/255: isolated one-level flips at .5 rounding boundaries are allowed."""
    from spaa_b200 import perc_al
    g = golden("extra")
    tiny = synth.TinyClassifier(3).to(dev())
    net = nn.Sequential(nn.AdaptiveAvgPool2d((20, 20)), tiny)
    imgs = synth.textured(121, "adv.x", (4, 3, 24, 32)).to(dev())
    true_lab, tgt_lab = T(g["adv_true"]).to(dev()), T(g["adv_tgt"]).to(dev())
    with torch.no_grad():
        order = net((imgs - 0.5) / 0.5).argsort(1, descending=True)
    assert torch.equal(order[:, 0], true_lab) and torch.equal(order[:, 2], tgt_lab)
    changed = 0
    for tag, labels, targeted, conf in (("t", tgt_lab, True, 0), ("u", true_lab, False, 0), ("u40", true_lab, False, 40)):
        atk = perc_al.PerC_AL(device=dev(), max_iterations=12, alpha_l_init=1, alpha_c_init=0.5, confidence=conf)
        out = atk.adversary(net, imgs.clone(), labels, targeted)
        ref = T(g["adv_" + tag])
        diff = (out.cpu() - ref).abs()
        assert diff.max().item() <= 1.01 / 255 and (diff > 1e-6).float().mean().item() < 2e-3, (tag, diff.max().item(), (diff > 1e-6).float().mean().item())
        changed += int((ref != imgs.cpu()).any())
    assert changed >= 1, "no variant produced an adversarial image: the fixture does not exercise the best-so-far copy"
    with pytest.raises(ValueError):
        atk.adversary(net, imgs + 1.0, true_lab, False)


def test_bf16x3_split_precision_mode_meets_the_fp32_fixtures(golden):
    """precision='bf16x3' (tensor cores, three bf16 parts per value, fp32 accumulation in TMEM) against the SAME fixtures and tolerances as the exact
    fp32 mode: PCNet / CompenNet++ outputs and input gradients (models.npz), the training loss trajectory and parameters after 3 Adam steps (train.npz)."""
    from spaa_b200 import models, ops, train_network as tn
    g = golden("models")
    P = synth.pcnet_params(31, CAM_HW)
    m = models.set_precision(make_pcnet(P, CAM_HW), "bf16x3")
    prj = synth.textured(32, "pc.prj", (2, 3, *PRJ_HW)).to(dev()).requires_grad_(True)
    scene = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).expand(2, -1, -1, -1).to(dev())
    probe = ops.set_probe(lambda kind, spec: kind.endswith("_tc"))
    y = m(prj, scene)
    close(y, g["pcnet_y"], 1e-5, 0, "pcnet y (bf16x3)")
    cot = synth.randn(34, "pc.cot", y.shape).to(dev())
    (y * cot).sum().backward()
    n_tc = len(probe["events"])
    ops.set_probe(None)
    assert n_tc >= 40, f"only {n_tc} launches went through the tcgen05 kernels"
    close(prj.grad, g["pcnet_gprj"], 2e-5, 1e-4, "pcnet gprj (bf16x3)")
    check_param_grads(g, "pcnet", m)
    C = synth.compennet_pp_params(37)
    cm = models.set_precision(make_cpp(C, PRJ_HW), "bf16x3")
    cam = synth.textured(38, "cpp.cam", (2, 3, *CAM_HW)).to(dev()).requires_grad_(True)
    yc = cm(cam, scene)
    close(yc, g["cpp_y"], 1e-5, 0, "cpp y (bf16x3)")
    (yc * synth.randn(39, "cpp.cot", yc.shape).to(dev())).sum().backward()
    close(cam.grad, g["cpp_gcam"], 2e-5, 1e-4, "cpp gcam (bf16x3)")
    # In THIS fixture two activations of the surface branch's last layer (r4s = relu(conv4_s(.)), the same pixel of both samples: the scene is shared)
    # have the pre-activation +7.45e-9 in the reference / the CUDA-core fp32 mode and <= 0 here (tools/diag_x3_grads.py: 2 sign mismatches of 32 768,
    # every other activation agrees to 7e-6): an exact tie at the ReLU threshold.  The gradients that pass through that mask -- the surface branch's
    # parameters and, through s = warp(scene), the warping net's -- differ by up to 6e-3 of their largest entry for that one reason; all other
    # parameters are held to the exact mode's 2e-5.
    check_param_grads(g, "cpp", cm, loose=(lambda n: "_s." in n or n.startswith("warping_net."), 1e-2))
    # training trajectory
    gt = golden("train")
    N = 6
    Pt = synth.pcnet_params(81, CAM_HW)
    mt = nn.DataParallel(models.set_precision(make_pcnet(Pt, CAM_HW), "bf16x3"), device_ids=[0])
    cfg = tn.AttrDict(device="cuda:0", data_root=None, setup_name="synth", model_name="PCNet", num_train=N, batch_size=4, max_iters=3, lr=1e-3,
                      lr_drop_ratio=0.2, lr_drop_rate=800, l2_reg=1e-4, plot_on=False, train_plot_rate=50, valid_rate=200, loss="l1+ssim")
    random.seed(5)
    tn.train_pcnet(mt, dict(cam_scene=synth.textured(83, "tr.scene", (1, 3, *CAM_HW)), cam_train=synth.textured(84, "tr.cam", (N, 3, *CAM_HW)),
                            prj_train=synth.textured(82, "tr.prj", (N, 3, *PRJ_HW)), mask=Pt["mask"]), None, cfg, verbose=False)
    close(cfg["loss_history"][:, 0], gt["pcnet_losses"], 1e-4, 0, "pcnet losses (bf16x3)")
    _check_after(gt, "pcnet", mt.module, Pt, 3e-4)
