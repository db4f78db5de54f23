"""CPU-side checks of the C-ABI boundary (no kernel is launched): libspaa_b200.so loads, exports every symbol that
include/spaa_b200.h declares, host-only entry points answer, argument errors come back as codes + text (never
exceptions across the ABI), and the host build of the per-pixel device math agrees with the oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

import synth
from oracle import spaa_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import importlib
    _lib = importlib.import_module("spaa_b200._lib")       # (spaa_b200._lib the attribute is rebound to the lib() accessor)
    protos = _lib.parse_header()
    assert len(protos) >= 35, sorted(protos)
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, f"declared in include/spaa_b200.h but not exported: {missing}"
    assert cdll.spaa_abi_version() >= 1
    # every exported spaa_* symbol is declared (no undocumented entry points)
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("spaa_")}
    assert exported == set(protos), exported ^ set(protos)


def test_host_only_entry_points_and_error_convention():
    from spaa_b200._lib import ConvDesc, SpaaError, lib
    L = lib()
    assert L.spaa_color_loss_ws_bytes(32, 240 * 320) > 0
    assert L.spaa_ssim_l1_ws_bytes(72, 240, 320) > 0
    assert L.spaa_rownorm_ws_bytes(32, 3 * 256 * 256) > 0
    d = ConvDesc()
    d.in_dtype = d.out_dtype = 1
    d.B, d.Cin, d.Hin, d.Win, d.Cout, d.Hout, d.Wout = 2, 128, 60, 80, 256, 60, 80
    d.KH = d.KW = 3
    d.stride = d.up = 1
    d.pad_h = d.pad_w = 1
    d.in_bs, d.in_ps, d.in_cs = 60 * 80 * 128, 128, 1
    d.out_bs, d.out_ps, d.out_cs = 60 * 80 * 256, 256, 1
    assert L.spaa_conv_tc_supported(ctypes.byref(d)) == 1
    assert L.spaa_conv_tc_packed_elems(ctypes.byref(d)) == 9 * 256 * 128
    d.in_dtype = 0                                      # fp32 activations are not a tensor-core case
    assert L.spaa_conv_tc_supported(ctypes.byref(d)) == 0
    d.in_dtype, d.Cin, d.in_ps, d.in_bs = 1, 48, 48, 60 * 80 * 48      # unsupported channel count
    assert L.spaa_conv_tc_supported(ctypes.byref(d)) == 0
    # argument errors: negative return code, message retrievable, surfaced as SpaaError by the binding (no launch happens)
    rc = L.cdll.spaa_rgb2lab_fwd(None, None, 1, 16, 0, None)
    assert rc < 0 and len(L.cdll.spaa_last_error()) > 0
    with pytest.raises(SpaaError):
        L.spaa_rgb2lab_fwd(None, None, 1, 16, 0, None)


def test_product_refuses_cpu_tensors():
    from spaa_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.rgb2lab(torch.rand(1, 3, 4, 4))


@pytest.fixture(scope="module")
def hostsim():
    path = os.path.join(ROOT, "tests", "hostsim", "_hostsim.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", path, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")], check=True)
    return ctypes.CDLL(path)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_device_colour_math_host_build_vs_oracle(hostsim):
    """The same color_math.cuh the kernels compile, built for the host: sRGB->Lab, the reference's dE2000 variant and the
    hand-derived backward formulas against the oracle's autograd (fp64: formulas; fp32: the arithmetic the GPU runs)."""
    n = 4096
    rgb1 = synth.textured(5, "hs.a", (1, 3, 64, 64)).reshape(3, n).double()
    rgb2 = (rgb1 + 0.05 * synth.randn(6, "hs.b", (3, n)).double()).clamp(0, 1)
    for dt, np_dt, sfx, tol in ((torch.float64, np.float64, "f64", 1e-9), (torch.float32, np.float32, "f32", 2e-4)):
        a = rgb1.to(dt).reshape(1, 3, 64, 64).requires_grad_(True)
        b = rgb2.to(dt).reshape(1, 3, 64, 64).requires_grad_(True)
        la, lb = O.srgb_to_lab(a), O.srgb_to_lab(b)
        de = O.de2000_variant(la, lb)
        a_np = np.ascontiguousarray(a.detach().numpy().reshape(3, n))
        lab = np.empty_like(a_np)
        getattr(hostsim, "hs_lab_fwd_" + sfx)(_ptr(a_np), _ptr(lab), ctypes.c_int64(n))
        assert np.abs(lab - la.detach().numpy().reshape(3, n)).max() <= tol * 100
        la_np = np.ascontiguousarray(la.detach().numpy().reshape(3, n))
        lb_np = np.ascontiguousarray(lb.detach().numpy().reshape(3, n))
        de_np = np.empty(n, np_dt)
        g1, g2 = np.empty_like(la_np), np.empty_like(la_np)
        getattr(hostsim, "hs_de_" + sfx)(_ptr(la_np), _ptr(lb_np), _ptr(de_np), _ptr(g1), _ptr(g2), ctypes.c_int64(n))
        assert np.abs(de_np - de.detach().numpy().reshape(n)).max() <= tol * 100
        if dt == torch.float64:            # derivative formulas: exact check in double precision
            ga, gb = torch.autograd.grad(de.sum(), (la, lb))
            assert np.abs(g1 - ga.numpy().reshape(3, n)).max() <= 1e-7
            assert np.abs(g2 - gb.numpy().reshape(3, n)).max() <= 1e-7


def test_device_warp_math_host_build_vs_oracle(hostsim):
    """warp_math.cuh (affine o TPS coarse grid and its parameter gradients) on the host against the oracle."""
    P = synth.warping_params(21)
    aff, theta, ctrl = (P["warping_net." + k].double() for k in ("affine_mat", "theta", "ctrl_pts"))
    aff.requires_grad_(True); theta.requires_grad_(True)
    Hin, Win, H, W = 32, 32, 24, 32
    a = O.affine_base_grid(aff, Hin, Win).permute(0, 3, 1, 2)
    t = O.tps_sampling_grid(theta, ctrl, H, W)
    ref = torch.nn.functional.grid_sample(a, t, align_corners=True).permute(0, 2, 3, 1)[0]      # H W 2
    T = ctrl.shape[0]
    out = np.empty((H, W, 2), np.float64)
    c = lambda x: np.ascontiguousarray(x.detach().numpy())
    an, tn, cn = c(aff), c(theta), c(ctrl)
    hostsim.hs_coarse_grid_f64(_ptr(an), _ptr(tn), _ptr(cn), T, Hin, Win, H, W, _ptr(out))
    assert np.abs(out - ref.detach().numpy()).max() <= 1e-10
    cot = synth.randn(7, "hs.cot", (H, W, 2)).double()
    ga, gt = torch.autograd.grad((ref * cot).sum(), (aff, theta))
    daff, dth = np.empty(6, np.float64), np.empty((T + 2) * 2, np.float64)
    hostsim.hs_coarse_grid_bwd_f64(_ptr(an), _ptr(tn), _ptr(cn), T, Hin, Win, H, W, _ptr(c(cot)), _ptr(daff), _ptr(dth))
    assert np.abs(daff - ga.numpy().reshape(-1)).max() <= 1e-8
    assert np.abs(dth - gt.numpy().reshape(-1)).max() <= 1e-8
