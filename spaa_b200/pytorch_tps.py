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


def _regular_xy(H: int, W: int, device):
    xs = torch.linspace(0, 1, W, device=device)
    ys = torch.linspace(0, 1, H, device=device)
    return xs, ys


def _tps_grid_one(theta, ctrl, H: int, W: int):
    """One TPS (theta (T+2)x2 reduced or (T+3)x2 full form, ctrl Tx2) -> planar sampling grid [2,H,W] in [-1,1]."""
    T = ctrl.shape[0]
    if theta.shape[0] == T + 2:
        return _TpsGridFn.apply(theta.unsqueeze(0), ctrl, H, W)
    if theta.shape[0] != T + 3:
        raise ValueError(f"theta must have T+2 (reduced) or T+3 (full) rows for T = {T} control points; got {theta.shape[0]}")
    # Full form (pytorch_tps.py:60-69 without the `reduced` branch): T free radial weights.  The kernel evaluates the reduced form, whose first
    # weight is tied to -sum(w_1..w_{T-1}); the difference to the full form is ONE radial term: (sum_t w_t) * U(|p - ctrl_0|).
    w, a = theta[:-3], theta[-3:]
    g = _TpsGridFn.apply(torch.cat((w[1:], a), 0).unsqueeze(0), ctrl, H, W)
    xs, ys = _regular_xy(H, W, theta.device)
    D = torch.sqrt((xs - ctrl[0, 0]).pow(2).unsqueeze(0) + (ys - ctrl[0, 1]).pow(2).unsqueeze(1))
    U0 = D.pow(2) * torch.log(D + 1e-6)
    return g + 2.0 * w.sum(0).view(2, 1, 1) * U0             # the grid is (xy + z) * 2 - 1: z enters with a factor 2


def _check(theta, ctrl):
    ops._need_cuda(theta, ctrl)
    if theta.dim() != 3 or theta.shape[2] != 2 or ctrl.dim() not in (2, 3) or (ctrl.dim() == 3 and ctrl.shape[0] != theta.shape[0]):
        raise ValueError("theta must be Nx(T+2|T+3)x2 and ctrl Tx2 or NxTx2")


def tps_grid(theta, ctrl, size):
    """pytorch_tps.py:79-106.  theta Nx(T+2)x2 (reduced) or Nx(T+3)x2 (full form), ctrl Tx2 or NxTx2, size (N,C,H,W) -> NxHxWx2 sampling grid in
    [-1,1].  One fused kernel launch per TPS of the batch (WarpingNet uses N = 1)."""
    _check(theta, ctrl)
    N, _, H, W = size
    if N != theta.shape[0]:
        raise ValueError(f"size[0] = {N} but theta holds {theta.shape[0]} parameter sets")
    gs = [_tps_grid_one(theta[n], ctrl if ctrl.dim() == 2 else ctrl[n], int(H), int(W)) for n in range(N)]
    return torch.stack(gs, 0).permute(0, 2, 3, 1)


def tps(theta, ctrl, grid):
    """pytorch_tps.py:29-76: TPS offsets z at the locations of a REGULAR NxHxWx3 grid (homogeneous 1, x = linspace(0,1,W), y = linspace(0,1,H): the
    grid tps_grid builds, :98-103, and the only one the reference ever passes).  Other sample locations raise: the fused kernel generates the regular
    grid analytically instead of reading it."""
    _check(theta, ctrl)
    N, H, W, _ = grid.shape
    xs, ys = _regular_xy(H, W, grid.device)
    if not (torch.allclose(grid[..., 1], xs.expand(N, H, W), atol=1e-6) and torch.allclose(grid[..., 2], ys.view(1, H, 1).expand(N, H, W), atol=1e-6)):
        raise NotImplementedError("spaa_b200.pytorch_tps.tps evaluates the TPS on the regular grid tps_grid builds (pytorch_tps.py:98-103)")
    full = tps_grid(theta, ctrl, (N, 1, H, W))
    return (full + 1) / 2 - grid[..., 1:]
