"""CPU oracle for the SPAA hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-torch (CPU, fp32 or fp64) restatement of the reference algorithms on the
hot path named by BASELINE.json `north_star`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s baseline legs (`cpu_baseline`, `--impl reference`,
and the `torch_cuda_reference` side leg, which runs these same stock-PyTorch ops on the
GPU as the denominator of the north-star's ">= 10x PyTorch-CUDA" target) may import this
module; nothing under `spaa_b200/` does.  Tensors are created on the device of the inputs.

Pinning: every function here is checked against the UNMODIFIED reference, imported in
the build container by `tests/golden/make_golden.py`, through the fixtures committed in
`tests/golden/*.npz` (see tests/test_oracle_golden.py).  The reference ships no tests or
golden vectors of its own, so these generated fixtures are the pin.

All `file:line` citations are relative to /root/reference/src/python/.
Functions are written functionally over a flat parameter dict that uses the reference's
state-dict key names (without the DataParallel `module.` prefix), so a reference
checkpoint can be fed to the oracle directly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ---------------------------------------------------------------------------------------
# tensor helpers (img_proc.py:110-132)
# ---------------------------------------------------------------------------------------


def to_4d(x: Tensor) -> Tensor:
    """img_proc.py:110-114: prepend singleton dims until x is BxCxHxW."""
    while x.ndim < 4:
        x = x.unsqueeze(0)
    return x


def crop_center(x: Tensor, size: Sequence[int]) -> Tensor:
    """img_proc.py:126-132: offsets are int(round((h-th)/2)) -- Python banker's rounding."""
    h, w = x.shape[-2:]
    th, tw = int(size[0]), int(size[1])
    top = int(round((h - th) / 2.0))
    left = int(round((w - tw) / 2.0))
    return x[..., top:top + th, left:left + tw]


def area_resize(x: Tensor, size: Sequence[int]) -> Tensor:
    """img_proc.py:117-123 -> F.interpolate(mode='area') == adaptive average pooling.

    Restated explicitly: output cell o averages input rows floor(o*H/h) .. ceil((o+1)*H/h)-1.
    """
    x4 = to_4d(x)
    H, W = x4.shape[-2:]
    oh, ow = int(size[0]), int(size[1])

    def pool_matrix(n_in: int, n_out: int) -> Tensor:
        m = torch.zeros(n_out, n_in, dtype=x4.dtype, device=x4.device)
        for o in range(n_out):
            lo = (o * n_in) // n_out
            hi = -((-(o + 1) * n_in) // n_out)
            m[o, lo:hi] = 1.0 / (hi - lo)
        return m

    ph = pool_matrix(H, oh)
    pw = pool_matrix(W, ow)
    y = torch.einsum("oh,bchw,pw->bcop", ph, x4, pw)
    return y.reshape(x.shape[:-2] + (oh, ow))


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def classifier_preprocess(im: Tensor, crop_sz: Sequence[int], input_sz: Sequence[int]) -> Tensor:
    """classifier.py:55-59: uint8->float/255, centre crop, area resize, ImageNet normalise."""
    if im.dtype == torch.uint8:
        im = im.to(torch.float32) / 255
    x = area_resize(crop_center(to_4d(im), crop_sz), input_sz)
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def classify(model, im: Tensor, crop_sz, input_sz):
    """classifier.py:55-72: returns (logits with graph, sorted probabilities, sorted indices)."""
    logits = model(classifier_preprocess(im, crop_sz, input_sz))
    p = F.softmax(logits, dim=1).detach()
    p_sorted, idx = p.sort(descending=True)
    return logits, p_sorted, idx


# ---------------------------------------------------------------------------------------
# colour: sRGB -> XYZ -> Lab and the reference's CIEDE2000 variant
# (perc_al/differential_color_functions.py:12-190)
# ---------------------------------------------------------------------------------------

_RGB2XYZ = ((0.4124, 0.3576, 0.1805), (0.2126, 0.7152, 0.0722), (0.0193, 0.1192, 0.9504))
_WHITE = (95.0489, 100.0, 108.8840)


def _srgb_linear_x100(c: Tensor) -> Tensor:
    """differential_color_functions.py:16-20.  Both branches are evaluated and blended with
    0/1 float masks (so a NaN in the unselected power branch propagates, as in the reference)."""
    hi = (c > 0.0405).to(c.dtype)
    gamma = ((c + 0.055) / 1.055) ** 2.4
    lin = hi * gamma + (1 - hi) * (c / 12.92)
    return 100 * lin


def _lab_f(t: Tensor) -> Tensor:
    """differential_color_functions.py:27-36: exact zeros are nudged by 1e-4, evaluated and
    then multiplied by 0, so f(0) = 0 (which makes black L = -16)."""
    is0 = (t == 0).to(t.dtype)
    t = t + 0.0001 * is0
    big = (t > 0.008856).to(t.dtype)
    f = big * t ** (1 / 3) + (1 - big) * (7.787 * t + 16 / 116)
    return f * (1 - is0)


def srgb_to_lab(rgb: Tensor) -> Tensor:
    """differential_color_functions.py:39-64 (rgb2lab_diff).  rgb: Bx3xHxW -> Lab Bx3xHxW."""
    lin = _srgb_linear_x100(rgb)
    m = torch.tensor(_RGB2XYZ, dtype=rgb.dtype, device=rgb.device)
    # :22 is a [3,3]x[3,BHW] matmul; einsum over the channel axis is the same contraction
    xyz = torch.einsum("kc,bchw->bkhw", m, lin)
    fx = _lab_f(xyz[:, 0] / _WHITE[0])
    fy = _lab_f(xyz[:, 1] / _WHITE[1])
    fz = _lab_f(xyz[:, 2] / _WHITE[2])
    return torch.stack((116 * fy - 16, 500 * (fx - fy), 200 * (fy - fz)), dim=1)


def _deg(r):
    return r * (180.0 / math.pi)


def _rad(d):
    return d * (math.pi / 180.0)


def _hue_deg(b: Tensor, a: Tensor) -> Tensor:
    """differential_color_functions.py:73-81 (hpf_diff): atan2 in degrees mapped to [0,360)."""
    both0 = ((b == 0) * (a == 0)).to(b.dtype)
    h = _deg(torch.atan2(b * (1 - both0), a * (1 - both0)))
    return h * (h >= 0).to(b.dtype) + (360 + h) * (h < 0).to(b.dtype)


def _hue_delta(c1, c2, h1, h2):
    """differential_color_functions.py:84-91 (dhpf_diff)."""
    nz = 1 - ((c1 * c2) == 0).to(c1.dtype)
    d = h2 - h1
    return (d * nz * (d.abs() <= 180).to(d.dtype)
            + (d - 360) * (d > 180).to(d.dtype) * nz
            + (d + 360) * (d < -180).to(d.dtype) * nz)


def _hue_mean(c1, c2, h1, h2):
    """differential_color_functions.py:94-106 (ahpf_diff), including the `+ (..)*mask1` term."""
    z = ((c1 * c2) == 0).to(c1.dtype)
    nz = 1 - z
    near = ((h2 - h1).abs() <= 180).to(c1.dtype)
    lt360 = ((h2 + h1).abs() < 360).to(c1.dtype)
    s = h1 + h2
    r = s * nz * near + (s + 360.0) * nz * (1 - near) * lt360 + (s - 360.0) * nz * (1 - near) * (1 - lt360)
    return (r + r * z) * 0.5


def de2000_variant(lab1: Tensor, lab2: Tensor) -> Tensor:
    """differential_color_functions.py:109-180 (ciede2000_diff).  Bx3xHxW x2 -> BxHxW.

    NOT textbook CIEDE2000: T uses 39 deg (:160), neutral inputs zero the chroma/hue/rotation
    terms (:128-133,147-155,172-173), and non-positive squares return 0 (:174-178).
    """
    L1, A1, B1 = lab1[:, 0], lab1[:, 1], lab1[:, 2]
    L2, A2, B2 = lab2[:, 0], lab2[:, 1], lab2[:, 2]
    dt = lab1.dtype
    n1 = ((A1 == 0) * (B1 == 0)).to(dt)
    n2 = ((A2 == 0) * (B2 == 0)).to(dt)
    B1 = B1 + 0.0001 * n1
    B2 = B2 + 0.0001 * n2
    C1 = torch.sqrt(A1 ** 2.0 + B1 ** 2.0)
    C2 = torch.sqrt(A2 ** 2.0 + B2 ** 2.0)
    cbar = (C1 + C2) / 2.0
    G = 0.5 * (1.0 - torch.sqrt(cbar ** 7.0 / (cbar ** 7.0 + 25 ** 7.0)))
    a1p = (1.0 + G) * A1
    a2p = (1.0 + G) * A2
    c1p = torch.sqrt(a1p ** 2.0 + B1 ** 2.0)
    c2p = torch.sqrt(a2p ** 2.0 + B2 ** 2.0)
    h1p = _hue_deg(B1, a1p) * (1 - n1)
    h2p = _hue_deg(B2, a2p) * (1 - n2)
    dLp = L2 - L1
    dCp = c2p - c1p
    dhp = _hue_delta(C1, C2, h1p, h2p)
    chroma_on = 1 - torch.max(n1, n2)
    dHp = 2.0 * torch.sqrt(c1p * c2p) * torch.sin(_rad(dhp) / 2.0) * chroma_on
    Lbar = (L1 + L2) / 2.0
    cpbar = (c1p + c2p) / 2.0
    hbar = _hue_mean(C1, C2, h1p, h2p)
    T = (1.0 - 0.17 * torch.cos(_rad(hbar - 39)) + 0.24 * torch.cos(_rad(2.0 * hbar))
         + 0.32 * torch.cos(_rad(3.0 * hbar + 6.0)) - 0.2 * torch.cos(_rad(4.0 * hbar - 63.0)))
    dtheta = 30.0 * torch.exp(-1.0 * ((hbar - 275.0) / 25.0) ** 2.0)
    rC = torch.sqrt(cpbar ** 7.0 / (cpbar ** 7.0 + 25.0 ** 7.0))
    sL = 1.0 + (0.015 * (Lbar - 50.0) ** 2.0) / torch.sqrt(20.0 + (Lbar - 50.0) ** 2.0)
    sC = 1.0 + 0.045 * cpbar
    sH = 1.0 + 0.015 * cpbar * T
    rT = -2.0 * rC * torch.sin(_rad(2.0 * dtheta))
    tl, tc, th = dLp / sL, dCp / sC, dHp / sH
    sq = tl ** 2.0 + tc ** 2.0 * chroma_on + th ** 2.0 * chroma_on + rT * tc * th * chroma_on
    nonpos = (sq <= 0).to(dt)
    return torch.sqrt(sq + 0.0001 * nonpos) * (1 - nonpos)


def mean_delta_e(x: Tensor, y: Tensor) -> float:
    """differential_color_functions.py:183-190 (deltaE): mean over batch and pixels, no grad."""
    with torch.no_grad():
        return de2000_variant(srgb_to_lab(to_4d(x)), srgb_to_lab(to_4d(y))).mean().item()


# ---------------------------------------------------------------------------------------
# thin-plate-spline grid (pytorch_tps.py:29-106, 201-217) and the warping grid
# ---------------------------------------------------------------------------------------


def uniform_ctrl(shape: Sequence[int]) -> Tensor:
    """pytorch_tps.py:201-217: control points on a regular lattice over [0,1]^2, (x,y) order."""
    gh, gw = int(shape[0]), int(shape[1])
    ys, xs = torch.meshgrid(torch.linspace(0, 1, gh), torch.linspace(0, 1, gw), indexing="ij")
    return torch.stack((xs, ys), dim=-1)


def tps_sampling_grid(theta: Tensor, ctrl: Tensor, H: int, W: int) -> Tensor:
    """pytorch_tps.py:79-106 + 54-74.  theta: 1x(T+2)x2 (reduced form) or 1x(T+3)x2;
    ctrl: Tx2 in [0,1].  Returns 1xHxWx2 sampling grid in [-1,1] (x,y order)."""
    dt = theta.dtype
    dv = theta.device
    ys, xs = torch.meshgrid(torch.linspace(0, 1, H, dtype=dt, device=dv), torch.linspace(0, 1, W, dtype=dt, device=dv), indexing="ij")
    xy = torch.stack((xs, ys), dim=-1)                                    # H W 2
    d = torch.sqrt(((xy.unsqueeze(-2) - ctrl.to(dv, dt)) ** 2).sum(-1))       # H W T
    U = d ** 2 * torch.log(d + 1e-6)
    w, a = theta[0, :-3], theta[0, -3:]
    if theta.shape[1] == ctrl.shape[0] + 2:                               # reduced form, :66-69
        w = torch.cat((-w.sum(0, keepdim=True), w), 0)
    ones = torch.ones(H, W, 1, dtype=dt, device=dv)
    z = torch.cat((ones, xy), -1) @ a + U @ w                             # H W 2
    return ((xy + z) * 2 - 1).unsqueeze(0)


def affine_base_grid(theta: Tensor, H: int, W: int) -> Tensor:
    """F.affine_grid(theta[1,2,3], (1,C,H,W), align_corners=True) (models.py:151,168) -> 1xHxWx2."""
    dt = theta.dtype
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, H, dtype=dt, device=theta.device), torch.linspace(-1, 1, W, dtype=dt, device=theta.device), indexing="ij")
    base = torch.stack((xs, ys, torch.ones_like(xs)), -1)                 # H W 3
    return (base @ theta[0].t()).unsqueeze(0)


def bilinear_sample(img: Tensor, grid: Tensor) -> Tensor:
    """F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    (models.py:155,172,184), restated with explicit gathers.  img BxCxHxW, grid BxhxWx2."""
    B, C, H, W = img.shape
    gx, gy = grid[..., 0], grid[..., 1]
    ix = (gx + 1) / 2 * (W - 1)
    iy = (gy + 1) / 2 * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    out = 0
    flat = img.reshape(B, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xc, yc = x0 + dx, y0 + dy
            wgt = (1 - (ix - xc).abs()) * (1 - (iy - yc).abs())
            ok = ((xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)).to(img.dtype)
            lin = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).long().reshape(B, 1, -1).expand(B, C, -1)
            v = flat.gather(2, lin).reshape(B, C, *gx.shape[1:])
            out = out + v * (wgt * ok).unsqueeze(1)
    return out


def _refine_net(P: Dict[str, Tensor], pre: str, g: Tensor) -> Tensor:
    """models.py:130-139: conv(2,32,3,s2,p1) ReLU conv(32,64,3,s2,p1) ReLU convT(64,32,2,s2) ReLU
    convT(32,2,2,s2) LeakyReLU(0.1)."""
    k = pre + "grid_refine_net."
    g = F.relu(F.conv2d(g, P[k + "0.weight"], P[k + "0.bias"], 2, 1))
    g = F.relu(F.conv2d(g, P[k + "2.weight"], P[k + "2.bias"], 2, 1))
    g = F.relu(F.conv_transpose2d(g, P[k + "4.weight"], P[k + "4.bias"], 2, 0))
    return F.leaky_relu(F.conv_transpose2d(g, P[k + "6.weight"], P[k + "6.bias"], 2, 0), 0.1)


def warping_fine_grid(P: Dict[str, Tensor], in_hw: Sequence[int], out_hw: Sequence[int],
                      pre: str = "warping_net.", with_refine: bool = True) -> Tensor:
    """models.py:149-178: affine grid at the INPUT size, TPS grid at out_size, the TPS grid samples
    the affine grid, optional refinement, clamp to [-1,1].  Returns 1 x h x w x 2."""
    aff = affine_base_grid(P[pre + "affine_mat"], in_hw[0], in_hw[1]).permute(0, 3, 1, 2)
    tps = tps_sampling_grid(P[pre + "theta"], P[pre + "ctrl_pts"], out_hw[0], out_hw[1])
    g = F.grid_sample(aff, tps, align_corners=True)
    if with_refine:
        g = _refine_net(P, pre, g) + g
    return torch.clamp(g, -1, 1).permute(0, 2, 3, 1)


def warp(P, x: Tensor, out_hw, pre="warping_net.", with_refine=True) -> Tensor:
    """models.py:163-185 (WarpingNet.forward, un-simplified)."""
    grid = warping_fine_grid(P, x.shape[2:], out_hw, pre, with_refine)
    return F.grid_sample(x, grid.expand(x.shape[0], -1, -1, -1), align_corners=True)


# ---------------------------------------------------------------------------------------
# ShadingNetSPAA / CompenNet / PCNet / CompenNet++ (models.py:11-94, 188-346)
# ---------------------------------------------------------------------------------------


def _cv(P, name, x, stride=1, pad=1):
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], stride, pad)


def _skip1(P, pre, s, first_pad):
    k = pre + "skipConv1."
    s = F.relu(F.conv2d(s, P[k + "0.weight"], P[k + "0.bias"], 1, first_pad))
    s = F.relu(F.conv2d(s, P[k + "2.weight"], P[k + "2.bias"], 1, 1))
    return F.relu(F.conv2d(s, P[k + "4.weight"], P[k + "4.bias"], 1, 1))


def shading_net(P, x: Tensor, *surf: Tensor, pre: str = "shading_net.", trace: Optional[dict] = None) -> Tensor:
    """models.py:280-303 (ShadingNetSPAA.forward).  `surf` = (s,) or (s, x*s); the skip branch
    runs on surf[0] (the surface image), not on x (:290-291)."""
    s = torch.cat(surf, 1)
    r1s = F.relu(_cv(P, pre + "conv1_s", s, 2))
    r2s = F.relu(_cv(P, pre + "conv2_s", r1s, 2))
    r3s = F.relu(_cv(P, pre + "conv3_s", r2s))
    r4s = F.relu(_cv(P, pre + "conv4_s", r3s))
    res1 = _skip1(P, pre, surf[0], 0)                     # first layer is 1x1 (:243)
    x1 = F.relu(_cv(P, pre + "conv1", x, 2) + r1s)
    res2 = _cv(P, pre + "skipConv2", x1, 1, 0)
    x2 = F.relu(_cv(P, pre + "conv2", x1, 2) + r2s)
    res3 = _cv(P, pre + "skipConv3", x2, 1, 1)            # 3x3 in ShadingNetSPAA (:252)
    x3 = F.relu(_cv(P, pre + "conv3", x2) + r3s)
    x4 = F.relu(_cv(P, pre + "conv4", x3) + r4s)
    x5 = F.relu(_cv(P, pre + "conv5", x4) + res3)
    x6 = F.relu(F.conv_transpose2d(x5, P[pre + "transConv1.weight"], P[pre + "transConv1.bias"], 2, 1, 1) + res2)
    x7 = F.relu(F.conv_transpose2d(x6, P[pre + "transConv2.weight"], P[pre + "transConv2.bias"], 2, 0))
    out = torch.clamp(F.relu(_cv(P, pre + "conv6", x7) + res1), max=1)
    if trace is not None:
        trace.update(r1s=r1s, r2s=r2s, r3s=r3s, r4s=r4s, res1=res1, x1=x1, res2=res2, x2=x2, res3=res3,
                     x3=x3, x4=x4, x5=x5, x6=x6, x7=x7)
    return out


def compen_net(P, x: Tensor, s: Tensor, pre: str = "compen_net.") -> Tensor:
    """models.py:74-94 (CompenNet.forward): 3-channel surface branch, 3x3 first skip layer run on x,
    1x1 skipConv3, both transposed convs k2 s2."""
    r1s = F.relu(_cv(P, pre + "conv1_s", s, 2))
    r2s = F.relu(_cv(P, pre + "conv2_s", r1s, 2))
    r3s = F.relu(_cv(P, pre + "conv3_s", r2s))
    r4s = F.relu(_cv(P, pre + "conv4_s", r3s))
    res1 = _skip1(P, pre, x, 1)
    x1 = F.relu(_cv(P, pre + "conv1", x, 2) + r1s)
    res2 = _cv(P, pre + "skipConv2", x1, 1, 0)
    x2 = F.relu(_cv(P, pre + "conv2", x1, 2) + r2s)
    res3 = _cv(P, pre + "skipConv3", x2, 1, 0)
    x3 = F.relu(_cv(P, pre + "conv3", x2) + r3s)
    x4 = F.relu(_cv(P, pre + "conv4", x3) + r4s)
    x5 = F.relu(_cv(P, pre + "conv5", x4) + res3)
    x6 = F.relu(F.conv_transpose2d(x5, P[pre + "transConv1.weight"], P[pre + "transConv1.bias"], 2, 0) + res2)
    x7 = F.relu(F.conv_transpose2d(x6, P[pre + "transConv2.weight"], P[pre + "transConv2.bias"], 2, 0))
    return torch.clamp(F.relu(_cv(P, pre + "conv6", x7) + res1), max=1)


def pcnet(P, prj: Tensor, scene: Tensor, out_hw, use_mask=True, use_rough=True, with_refine=True,
          trace: Optional[dict] = None) -> Tensor:
    """models.py:335-346 (PCNet.forward): warp -> x mask -> ShadingNet(x, s, x*s)."""
    x = warp(P, prj, out_hw, "warping_net.", with_refine)
    if use_mask:
        x = x * P["mask"].to(x.dtype)
    if trace is not None:
        trace["warped"] = x
    if use_rough:
        return shading_net(P, x, scene, x * scene, trace=trace)
    return shading_net(P, x, scene, trace=trace)


def compennet_pp(P, cam: Tensor, scene: Tensor, out_hw, with_refine=True) -> Tensor:
    """models.py:204-212 (CompenNetPlusplus.forward): warp BOTH x and s, then CompenNet."""
    return compen_net(P, warp(P, cam, out_hw, "warping_net.", with_refine),
                      warp(P, scene, out_hw, "warping_net.", with_refine))


# ---------------------------------------------------------------------------------------
# SSIM + training loss (pytorch_ssim/__init__.py:9-107, train_network.py:367-392)
# ---------------------------------------------------------------------------------------


def gauss_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> Tensor:
    """pytorch_ssim/__init__.py:9-12: normalised 1-D Gaussian (built in fp32 like torch.Tensor)."""
    g = torch.tensor([math.exp(-(i - size // 2) ** 2 / float(2 * sigma ** 2)) for i in range(size)],
                     dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def ssim_map(a: Tensor, b: Tensor, size: int = 11) -> Tensor:
    """pytorch_ssim/__init__.py:24-51: replicate-pad, 11x11 Gaussian moments, C1=1e-4, C2=9e-4."""
    ch = a.shape[1]
    g = gauss_window(size, dtype=torch.float32)
    win = (g[:, None] @ g[None, :]).to(a.device, a.dtype).expand(ch, 1, size, size).contiguous()
    p = size // 2
    a = F.pad(a, (p, p, p, p), mode="replicate")
    b = F.pad(b, (p, p, p, p), mode="replicate")
    mu_a = F.conv2d(a, win, groups=ch)
    mu_b = F.conv2d(b, win, groups=ch)
    va = F.conv2d(a * a, win, groups=ch) - mu_a.pow(2)
    vb = F.conv2d(b * b, win, groups=ch) - mu_b.pow(2)
    cab = F.conv2d(a * b, win, groups=ch) - mu_a * mu_b
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu_a * mu_b + C1) * (2 * cab + C2)) / ((mu_a.pow(2) + mu_b.pow(2) + C1) * (va + vb + C2))


def ssim_index(a: Tensor, b: Tensor, size_average: bool = True, mask=None, weights=None) -> Tensor:
    """pytorch_ssim/__init__.py:53-67."""
    m = ssim_map(a, b)
    if weights is not None:
        m = m * weights.expand_as(m)
    if size_average:
        return m[mask].mean() if mask is not None else m.mean()
    if mask is not None:
        return (m * mask).mean(1).mean(1).mean(1)
    return m.mean(1).mean(1).mean(1)


def training_loss(pred: Tensor, target: Tensor, option: str) -> Tuple[Tensor, Tensor]:
    """train_network.py:367-392 (compute_loss): substring-selected terms; MSE always returned."""
    if option == "":
        raise TypeError("Loss type not specified")
    total = 0
    if "l1" in option:
        total = total + F.l1_loss(pred, target)
    mse = F.mse_loss(pred, target)
    if "l2" in option:
        total = total + mse
    if "ssim" in option:
        total = total + (1 - ssim_index(pred, target))
    if "huber" in option:
        d2 = (pred - target) ** 2
        total = total + (((1 + d2 / 0.1 ** 2).clamp(1e-4).sqrt() - 1) * 0.1).abs().mean()
    return total, mse


# ---------------------------------------------------------------------------------------
# SPAA attack loop (projector_based_attack.py:212-339)
# ---------------------------------------------------------------------------------------


def _per_sample_unit(g: Tensor) -> Tensor:
    """g / ||g||_2 with the norm over all C*H*W elements of each sample (no epsilon), :307,315."""
    n = torch.norm(g.reshape(g.shape[0], -1), dim=1)
    return g / n.view(-1, 1, 1, 1)


def spaa_attack(pcnet_fn, classifier_fn, target_idx: Sequence[int], targeted: bool, cam_scene: Tensor, d_thr: float,
                stealth_loss: str, prj_hw=(256, 256), prj_brightness: float = 0.5, iters: int = 50,
                trace: Optional[List[dict]] = None, forced_prj: Optional[List[Tensor]] = None):
    """projector_based_attack.py:212-339 (spaa).

    pcnet_fn(prj, scene) -> cam_infer;  classifier_fn(img) -> (logits, p_sorted, idx_sorted).
    `trace`, if given, receives one dict per iteration (inputs, losses, masks, both gradients and the
    post-update state) for teacher-forced parity tests.  `forced_prj[i]`, if given, overrides the
    projector image at the start of iteration i (teacher forcing).
    """
    B = len(target_idx)
    scene = to_4d(cam_scene)
    scene_b = scene.expand(B, -1, -1, -1)
    tgt = torch.as_tensor(list(target_idx), dtype=torch.long, device=scene.device)
    gray = prj_brightness * torch.ones(B, 3, *prj_hw, dtype=scene.dtype, device=scene.device)
    prj = gray.clone().requires_grad_(True)
    adv_lr, col_lr, p_thresh = 2, 1, 0.9                                  # :243-255
    w_prjl2 = 0.1 if "prjl2" in stealth_loss else 0
    w_caml2 = 1 if "caml2" in stealth_loss else 0
    w_camde = 1 if "camdE" in stealth_loss else 0
    best_prj = prj.detach().clone()
    best_cam = scene.repeat(B, 1, 1, 1).clone()
    best_col = 1e6 * torch.ones(B, dtype=scene.dtype, device=scene.device)
    ar = torch.arange(B, device=scene.device)

    for it in range(iters):
        if forced_prj is not None:
            prj.data.copy_(forced_prj[it])
        prj_in = prj.detach().clone()
        cam = pcnet_fn(torch.clamp(prj, 0, 1), scene_b)                   # :265
        logits, p, idx = classifier_fn(cam)                               # :266
        sel = logits[ar, tgt]
        adv_loss = (-sel).mean() if targeted else sel.mean()              # :269-272
        prjl2 = torch.norm(gray - prj, dim=1).mean(1).mean(1)             # :275
        caml2 = torch.norm(scene_b - cam, dim=1).mean(1).mean(1)          # :279
        camde = de2000_variant(srgb_to_lab(cam), srgb_to_lab(scene_b)).mean(1).mean(1)   # :283
        col_b = w_prjl2 * prjl2 + w_caml2 * caml2 + w_camde * camde
        col_loss = col_b.mean()
        high_conf = p[:, 0] > p_thresh                                    # :290
        high_pert = (caml2 * 255 > d_thr).detach()                        # :291
        if targeted:
            succ = idx[:, 0] == tgt
            use_col = succ & high_conf & high_pert                        # :295-296
        else:
            succ = idx[:, 0] != tgt
            use_col = succ & high_pert                                    # :298-299
        g_adv, = torch.autograd.grad(adv_loss, prj, retain_graph=True)    # :302-304
        prj.data[~use_col] -= adv_lr * _per_sample_unit(g_adv)[~use_col]  # :307
        g_col, = torch.autograd.grad(col_loss, prj)                       # :310-312
        prj.data[use_col] -= col_lr * _per_sample_unit(g_col)[use_col]    # :315
        better = (col_b.detach() < best_col) & use_col                    # :318-319
        best_col[better] = col_b.detach()[better]
        best_prj[succ] = prj.detach()[succ]                               # :323-324 (updated prj, pre-update cam)
        best_cam[succ] = cam.detach()[succ]
        best_prj[better] = prj.detach()[better]                           # :327-328
        best_cam[better] = cam.detach()[better]
        if trace is not None:
            trace.append(dict(prj_in=prj_in, cam=cam.detach().clone(), logits=logits.detach().clone(),
                              adv_loss=adv_loss.detach().clone(), caml2=caml2.detach().clone(),
                              camde=camde.detach().clone(), prjl2=prjl2.detach().clone(),
                              col_b=col_b.detach().clone(), use_col=use_col.clone(), succ=succ.clone(),
                              g_adv=g_adv.clone(), g_col=g_col.clone(), prj_out=prj.detach().clone(),
                              best_col=best_col.clone(), best_prj=best_prj.clone(), best_cam=best_cam.clone()))
    return best_cam, torch.clamp(best_prj, 0, 1)                          # :337-339


# ---------------------------------------------------------------------------------------
# PerC-AL (perc_al/__init__.py:133-256) + CompenNet++ (projector_based_attack.py:342-359)
# ---------------------------------------------------------------------------------------


def perc_al_attack(classifier_fn, inputs: Tensor, labels: Tensor, d_thr: float, targeted: bool,
                   max_iterations: int = 50, alpha_l_init: float = 1.0, alpha_c_init: float = 0.5,
                   confidence: float = 0, trace: Optional[List[dict]] = None) -> Tensor:
    """perc_al/__init__.py:133-256 (PerC_AL.adversary_projector)."""
    if inputs.min() < 0 or inputs.max() > 1:
        raise ValueError("Input values should be in the [0, 1] range.")
    a_l_min, a_c_min = alpha_l_init / 100, alpha_c_init / 10
    sign = -1 if targeted else 1
    B = inputs.shape[0]
    best = inputs.clone()
    lab0 = srgb_to_lab(inputs)
    delta = torch.zeros_like(inputs, requires_grad=True)
    use_col = torch.zeros(B, dtype=torch.bool, device=inputs.device)
    best_dis = torch.ones(B, dtype=inputs.dtype, device=inputs.device) * 100000
    if targeted and confidence != 0:
        return None                                                       # :176-178
    for it in range(max_iterations):
        logits, _, _ = classifier_fn(inputs + delta)                      # :181
        cosf = 1 + math.cos(it / max_iterations * math.pi)
        a_c = a_c_min + 0.5 * (alpha_c_init - a_c_min) * cosf             # :184-185
        a_l = a_l_min + 0.5 * (alpha_l_init - a_l_min) * cosf
        loss = sign * F.cross_entropy(logits, labels, reduction="sum")    # :186
        g_a, = torch.autograd.grad(loss, delta)
        delta.data[~use_col] = delta.data[~use_col] + a_l * _per_sample_unit(g_a)[~use_col]    # :193-195
        dmap = de2000_variant(lab0, srgb_to_lab(inputs + delta))          # :197
        dis = torch.norm(dmap.reshape(B, -1), dim=1)
        g_c, = torch.autograd.grad(dis.sum(), delta)
        delta.data[use_col] = delta.data[use_col] - a_c * _per_sample_unit(g_c)[use_col]       # :205-209
        delta.data = (inputs + delta.data).clamp(0, 1) - inputs           # :211
        x_round = torch.round((inputs + delta.data) * 255) / 255          # :212, :15-18
        caml2 = torch.norm(delta.detach(), dim=1).mean(1).mean(1)         # :215
        high_pert = caml2 * 255 > d_thr
        with torch.no_grad():
            logits2, p2, idx2 = classifier_fn(x_round)                    # :220/229/235
        high_conf = p2[:, 0] > 0.9
        if (not targeted) and confidence != 0:
            real = logits2.gather(1, labels.unsqueeze(1)).squeeze(1)
            inf_hot = torch.zeros_like(logits2).scatter_(1, labels.unsqueeze(1), float("inf"))
            other = (logits2 - inf_hot).max(1)[0]
            isadv = (real - other) <= -40                                 # :224 (literal 40, not `confidence`)
            use_col = isadv & high_pert
        elif targeted:
            isadv = idx2[:, 0] == labels
            use_col = isadv & high_conf & high_pert                       # :229-232
        else:
            isadv = idx2[:, 0] != labels
            use_col = isadv & high_pert                                   # :235-238
        better = (dis.detach() < best_dis) & use_col                      # :240-242
        best_dis[better] = dis.detach()[better]
        best[isadv] = x_round[isadv]                                      # :244-245
        best[better] = x_round[better]
        if trace is not None:
            trace.append(dict(g_a=g_a.clone(), g_c=g_c.clone(), delta=delta.detach().clone(), dis=dis.detach().clone(),
                              x_round=x_round.clone(), use_col=use_col.clone(), isadv=isadv.clone(), best=best.clone()))
    return best


def perc_al_compennet_pp_attack(compennet_pp_fn, classifier_fn, target_idx, targeted, cam_scene, d_thr, iters=50,
                                trace=None):
    """projector_based_attack.py:342-359."""
    B = len(target_idx)
    scene_b = to_4d(cam_scene).expand(B, -1, -1, -1)
    best_cam = perc_al_attack(classifier_fn, scene_b, torch.as_tensor(list(target_idx), dtype=torch.long), d_thr,
                              targeted, max_iterations=iters, alpha_l_init=1, alpha_c_init=0.5,
                              confidence=0 if targeted else 40, trace=trace)
    with torch.no_grad():
        return best_cam, compennet_pp_fn(best_cam, scene_b)


# ---------------------------------------------------------------------------------------
# PCNet / CompenNet++ training step (train_network.py:235-363, 130-232)
# ---------------------------------------------------------------------------------------


def pcnet_param_groups(names: Sequence[str]) -> Tuple[List[str], List[str], List[str]]:
    """train_network.py:248-250: (affine+theta | refine net | everything else)."""
    g1 = [n for n in names if n in ("warping_net.affine_mat", "warping_net.theta")]
    g2 = [n for n in names if "warping_net.grid_refine_net" in n]
    g3 = [n for n in names if "warping_net" not in n]
    return g1, g2, g3


def pcnet_loss_name(step: int) -> str:
    """train_network.py:300-303."""
    return "l1" if step <= 400 else "l1+ssim"


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float = 0.0,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (L2-in-gradient weight decay) as used at train_network.py:253-255,145; in place."""
    if wd != 0:
        g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def multistep_lr(base: float, step: int, milestones: Sequence[int], gamma: float) -> float:
    """optim.lr_scheduler.MultiStepLR as used at train_network.py:263-265 (step = #scheduler.step() calls)."""
    return base * gamma ** sum(1 for m in milestones if step >= m)
