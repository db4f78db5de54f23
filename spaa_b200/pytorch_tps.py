"""Thin-plate-spline sampling grids -- API of /root/reference/src/python/pytorch_tps.py:29-106, 201-217.

`tps_grid` / `tps` run the fused sm_100a kernel (spaa_tps_grid_fwd / spaa_coarse_grid_bwd): the T radial-basis terms are
evaluated per output pixel in registers; the reference's NxHxWxT `U` tensor and its two bmm calls never exist.
"""
from __future__ import annotations

import torch

from . import ops


def uniform_grid(shape):
    """pytorch_tps.py:201-217: HxWx2 control points over [0,1]^2, (x, y) order.  Host-side constant."""
    H, W = shape[:2]
    c = torch.zeros(H, W, 2)
    c[..., 0] = torch.linspace(0, 1, W)
    c[..., 1] = torch.linspace(0, 1, H).unsqueeze(-1)
    return c


class _TpsGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, ctrl, H, W):
        ctx.save_for_backward(theta, ctrl)
        ctx.hw = (H, W)
        return ops.tps_grid(theta, ctrl, H, W)            # planar [2,H,W]

    @staticmethod
    def backward(ctx, dgrid):
        theta, ctrl = ctx.saved_tensors
        _, dtheta = ops.coarse_grid_bwd(None, theta, ctrl, (0, 0), ctx.hw, dgrid)
        return dtheta.view_as(theta), None, None, None


def _check(theta, ctrl):
    ops._need_cuda(theta, ctrl)
    if theta.shape[0] != 1 or ctrl.dim() != 2:
        raise NotImplementedError("spaa_b200.pytorch_tps handles one TPS (N=1, ctrl Tx2), as WarpingNet uses it")
    if theta.shape[1] != ctrl.shape[0] + 2:
        raise NotImplementedError("only the reduced TPS form (T+2 parameter rows) is implemented")


def tps_grid(theta, ctrl, size):
    """pytorch_tps.py:79-106.  theta 1x(T+2)x2, ctrl Tx2, size (N,C,H,W) -> 1xHxWx2 sampling grid in [-1,1]."""
    _check(theta, ctrl)
    _, _, H, W = size
    g = _TpsGridFn.apply(theta, ctrl, int(H), int(W))
    return g.permute(1, 2, 0).unsqueeze(0)


def tps(theta, ctrl, grid):
    """pytorch_tps.py:29-76: TPS offsets z at the locations of a regular 1xHxWx3 grid (as built by tps_grid)."""
    _check(theta, ctrl)
    _, H, W, _ = grid.shape
    full = tps_grid(theta, ctrl, (1, 1, H, W))
    return (full + 1) / 2 - grid[..., 1:]
