"""Deterministic synthetic inputs shared by make_golden.py (runs the reference here), the
tests and bench.py (run anywhere).  Nothing here depends on /root/reference or on the order in
which a module constructor consumes torch's global RNG: every tensor is drawn from its own
generator seeded by (seed, crc32(key))."""
from __future__ import annotations

import zlib
from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def randn(seed: int, key: str, shape, scale: float = 1.0) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(seed, key)) * scale


def rand(seed: int, key: str, shape) -> torch.Tensor:
    return torch.rand(tuple(shape), generator=_gen(seed, key))


# ----------------------------------------------------------------------------------------------
# parameter tables: name -> (shape, fan_in) in the reference's state-dict naming
# (models.py:18-46, 131-137, 223-252)
# ----------------------------------------------------------------------------------------------

def _conv(cout, cin, k):
    return (cout, cin, k, k), cin * k * k, cout


def _convT(cin, cout, k):
    return (cin, cout, k, k), cin * k * k, cout


def shading_net_shapes(use_rough: bool = True) -> Dict[str, tuple]:
    ns = 6 if use_rough else 3
    t = {
        "conv1": _conv(32, 3, 3), "conv2": _conv(64, 32, 3), "conv3": _conv(128, 64, 3),
        "conv4": _conv(256, 128, 3), "conv5": _conv(128, 256, 3),
        "conv1_s": _conv(32, ns, 3), "conv2_s": _conv(64, 32, 3), "conv3_s": _conv(128, 64, 3),
        "conv4_s": _conv(256, 128, 3),
        "transConv1": _convT(128, 64, 3), "transConv2": _convT(64, 32, 2), "conv6": _conv(3, 32, 3),
        "skipConv1.0": _conv(3, 3, 1), "skipConv1.2": _conv(3, 3, 3), "skipConv1.4": _conv(3, 3, 3),
        "skipConv2": _conv(64, 32, 1), "skipConv3": _conv(128, 64, 3),
    }
    return t


def compen_net_shapes() -> Dict[str, tuple]:
    t = shading_net_shapes(False)
    t["transConv1"] = _convT(128, 64, 2)
    t["skipConv1.0"] = _conv(3, 3, 3)
    t["skipConv3"] = _conv(128, 64, 1)
    return t


def refine_net_shapes() -> Dict[str, tuple]:
    return {"grid_refine_net.0": _conv(32, 2, 3), "grid_refine_net.2": _conv(64, 32, 3),
            "grid_refine_net.4": _convT(64, 32, 2), "grid_refine_net.6": _convT(32, 2, 2)}


def _fill(seed, prefix, table, gain=1.0, bias_scale=0.05) -> Dict[str, torch.Tensor]:
    out = {}
    for name, (shape, fan_in, cout) in table.items():
        std = gain * (2.0 / fan_in) ** 0.5
        out[f"{prefix}{name}.weight"] = randn(seed, prefix + name + ".w", shape, std)
        out[f"{prefix}{name}.bias"] = randn(seed, prefix + name + ".b", (cout,), bias_scale)
    return out


def warping_params(seed: int, prefix: str = "warping_net.", grid_shape=(6, 6), refine_gain: float = 0.02,
                   theta_scale: float = 0.01, affine=(1.05, 0.03, 0.02, -0.04, 0.97, -0.01)) -> Dict[str, torch.Tensor]:
    """WarpingNet parameters + ctrl_pts buffer (models.py:114-121).  refine_gain is far above the
    reference's N(0,1e-4) init so the refinement net visibly moves the grid in parity tests."""
    nctrl = grid_shape[0] * grid_shape[1]
    p = {prefix + "affine_mat": torch.tensor(affine, dtype=torch.float32).view(1, 2, 3),
         prefix + "theta": randn(seed, prefix + "theta", (1, nctrl + 2, 2), theta_scale)}
    ys, xs = torch.meshgrid(torch.linspace(0, 1, grid_shape[0]), torch.linspace(0, 1, grid_shape[1]), indexing="ij")
    p[prefix + "ctrl_pts"] = torch.stack((xs, ys), -1).view(-1, 2)
    p.update(_fill(seed, prefix, refine_net_shapes(), gain=refine_gain, bias_scale=0.002))
    return p


def pcnet_params(seed: int, cam_hw: Sequence[int], gain: float = 0.7, use_rough: bool = True) -> Dict[str, torch.Tensor]:
    """Full PCNet state dict (no `module.` prefix).  gain<1 keeps the random net out of saturation
    (a kaiming-init ShadingNet clamps most outputs to 1)."""
    p = warping_params(seed)
    p.update(_fill(seed, "shading_net.", shading_net_shapes(use_rough), gain=gain))
    p["mask"] = quad_mask(cam_hw)
    return p


def compennet_pp_params(seed: int, gain: float = 0.7) -> Dict[str, torch.Tensor]:
    p = warping_params(seed, affine=(0.96, -0.02, 0.01, 0.03, 1.02, 0.02))
    p.update(_fill(seed, "compen_net.", compen_net_shapes(), gain=gain))
    return p


def quad_mask(hw: Sequence[int]) -> torch.Tensor:
    """Centred convex quadrilateral 0/1 float mask [1,1,H,W] (stands in for the direct-light mask)."""
    H, W = int(hw[0]), int(hw[1])
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    inside = (ys.abs() * 1.0 + xs.abs() * 0.12 < 0.93) & (xs.abs() * 1.0 + ys.abs() * 0.08 < 0.95)
    return inside.float().view(1, 1, H, W)


def textured(seed: int, key: str, shape, lo: float = 0.05, hi: float = 0.95) -> torch.Tensor:
    """Low-pass-filtered uniform noise at three spatial scales, rescaled to [lo,hi]; shape B x C x H x W."""
    B, C, H, W = shape
    acc = torch.zeros(shape)
    for i, div in enumerate((1, 4, 16)):
        h, w = max(H // div, 1), max(W // div, 1)
        n = rand(seed, f"{key}.{i}", (B, C, h, w))
        acc = acc + F.interpolate(n, size=(H, W), mode="bilinear", align_corners=False) * (0.5 + i)
    acc = acc - acc.amin(dim=(1, 2, 3), keepdim=True)
    acc = acc / acc.amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)
    return lo + (hi - lo) * acc


class TinyClassifier(nn.Module):
    """A small stand-in 1000-way classifier for CPU-sized parity runs (weights from synth.randn)."""

    def __init__(self, seed: int = 0, logit_scale: float = 8.0):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, 2, 1)
        self.c2 = nn.Conv2d(8, 16, 3, 2, 1)
        self.fc = nn.Linear(16 * 4 * 4, 1000)
        self.logit_scale = logit_scale
        with torch.no_grad():
            for n, p in self.named_parameters():
                fan = p[0].numel() if p.ndim > 1 else 1
                p.copy_(randn(seed, "tiny." + n, p.shape, (2.0 / fan) ** 0.5 if p.ndim > 1 else 0.1))
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        x = F.relu(self.c1(x))
        x = F.relu(self.c2(x))
        x = F.adaptive_avg_pool2d(x, 4).flatten(1)
        return self.fc(x) * self.logit_scale


SPAA_TARGETS10 = (1, 7, 21, 207, 340, 745, 779, 846, 947, 950)   # data/imagenet10_clsidx_to_labels.txt


def to_numpy_dict(d: Dict[str, torch.Tensor]):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}
