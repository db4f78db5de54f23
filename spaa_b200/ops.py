"""Thin tensor-level wrappers over the C ABI (include/spaa_b200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every function here launches hand-written
sm_100a kernels from libspaa_b200.so on torch's current CUDA stream.  There is no CPU path: inputs must be CUDA
tensors and the library must be built (spaa_b200.build) -- otherwise an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch

from ._lib import ConvDesc, lib

Tensor = torch.Tensor

EPI_RELU, EPI_LEAKY01, EPI_CLAMP_MAX1, EPI_ADD_AFTER_ACT, EPI_OUT2_BF16 = 1, 2, 4, 8, 16
MASK_NONE, MASK_POS, MASK_LEAKY01, MASK_OPEN01 = 0, 1, 2, 3

_launches = 0      # number of kernels launched through this module (bench.py reports it)


def launch_count() -> int:
    return _launches


_probe = None      # optional per-launch timing of one kernel family: {"match": fn(kind, spec), "events": [(start, end), ...]}


def set_probe(match=None):
    """bench.py: time every launch for which match(kind, spec) is true with a CUDA-event pair on the launching stream."""
    global _probe
    _probe = None if match is None else {"match": match, "events": []}
    return _probe


class _Probe:
    def __init__(self, kind, spec):
        self.on = _probe is not None and _probe["match"](kind, spec) and not torch.cuda.is_current_stream_capturing()

    def __enter__(self):
        if self.on:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            _probe["events"].append((self.e0, self.e1))


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def _p(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spaa_b200 kernels need CUDA tensors (there is no CPU fallback); got a "
                               f"{t.device} tensor of shape {tuple(t.shape)}")


def _f32c(t: Tensor) -> Tensor:
    _need_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


_ws_cache: Dict[Tuple[str, int, int], Tensor] = {}


_ws_scope_id = 0


class ws_scope:
    """Workspaces requested inside `with ws_scope(k):` are private to scope k.  An attack engine runs its iteration inside its own
    scope, so engines that execute concurrently on different streams (half-batch pipelining) never share a scratch buffer."""

    def __init__(self, scope: int):
        self.scope, self.prev = int(scope), 0

    def __enter__(self):
        global _ws_scope_id
        self.prev, _ws_scope_id = _ws_scope_id, self.scope
        return self

    def __exit__(self, *exc):
        global _ws_scope_id
        _ws_scope_id = self.prev
        return False


def release_scope(scope: int) -> None:
    """Drop the workspaces of a finished `ws_scope` (called when an attack engine is destroyed)."""
    suffix = f"@{int(scope)}"
    for k in [k for k in _ws_cache if k[0].endswith(suffix)]:
        del _ws_cache[k]


def workspace(tag: str, nbytes: int, device) -> Tensor:
    """Zero-initialised scratch owned by this module; kernels leave it zeroed (self-resetting counters), so it is
    reused across launches on the same stream.  Distinct `tag`s (and distinct `ws_scope`s) never alias."""
    dev = torch.device(device)
    key = (tag if _ws_scope_id == 0 else f"{tag}@{_ws_scope_id}", dev.index if dev.index is not None else torch.cuda.current_device(), int(nbytes))
    t = _ws_cache.get(key)
    if t is None:
        t = torch.zeros((int(nbytes) + 3) // 4, dtype=torch.int32, device=dev)
        _ws_cache[key] = t
    return t


# ------------------------------------------------------------------------------------------------------------
# colour
# ------------------------------------------------------------------------------------------------------------

def rgb2lab(rgb: Tensor, fast: bool = False) -> Tensor:
    rgb = _f32c(rgb)
    B, C, H, W = rgb.shape
    assert C == 3
    lab = torch.empty_like(rgb)
    lib().spaa_rgb2lab_fwd(_p(rgb), _p(lab), B, H * W, int(bool(fast)), _stream()); _count()
    return lab


def rgb2lab_bwd(rgb: Tensor, dlab: Tensor) -> Tensor:
    rgb, dlab = _f32c(rgb), _f32c(dlab)
    B, _, H, W = rgb.shape
    out = torch.empty_like(rgb)
    lib().spaa_rgb2lab_bwd(_p(rgb), _p(dlab), _p(out), B, H * W, _stream()); _count()
    return out


def _bstride(t: Tensor, B: int) -> int:
    if t.shape[0] == B:
        return t.stride(0)
    if t.shape[0] == 1:
        return 0
    raise ValueError(f"batch {t.shape[0]} does not broadcast to {B}")


def de2000(lab1: Tensor, lab2: Tensor) -> Tensor:
    lab1, lab2 = _f32c(lab1), _f32c(lab2)
    B = max(lab1.shape[0], lab2.shape[0])
    H, W = lab1.shape[-2:]
    de = torch.empty((B, H, W), dtype=torch.float32, device=lab1.device)
    lib().spaa_de2000_fwd(_p(lab1), _bstride(lab1, B), _p(lab2), _bstride(lab2, B), _p(de), B, H * W, _stream()); _count()
    return de


def de2000_bwd(lab1: Tensor, lab2: Tensor, cot: Tensor, need1: bool = True, need2: bool = True):
    lab1, lab2, cot = _f32c(lab1), _f32c(lab2), _f32c(cot)
    B = max(lab1.shape[0], lab2.shape[0])
    H, W = lab1.shape[-2:]
    d1 = torch.empty((B, 3, H, W), dtype=torch.float32, device=lab1.device) if need1 else None
    d2 = torch.empty((B, 3, H, W), dtype=torch.float32, device=lab1.device) if need2 else None
    lib().spaa_de2000_bwd(_p(lab1), _bstride(lab1, B), _p(lab2), _bstride(lab2, B), _p(cot), _p(d1), _p(d2), B, H * W,
                          _stream()); _count()
    if d1 is not None and lab1.shape[0] == 1 and B > 1:
        d1 = d1.sum(0, keepdim=True)
    if d2 is not None and lab2.shape[0] == 1 and B > 1:
        d2 = d2.sum(0, keepdim=True)
    return d1, d2


def color_loss(cam: Tensor, ref_rgb: Tensor, ref_lab: Tensor, *, cam_is_lab2: bool, de_weighting: bool, c_de: float,
               c_l2: float, stats: Optional[Tensor] = None, grad: Optional[Tensor] = None, want_grad: bool = True, fast: bool = False):
    """Fused Lab + dE2000 + channel-L2 statistics and gradient (spaa_color_loss_fwd_bwd).  fast: hardware-approximation arithmetic (opt-in beside a 16-bit PCNet; never the fp32
    parity mode); `ref_lab` must then be rgb2lab(ref_rgb, fast=True).
    Returns (stats [B,4] = sum dE, sum ||.||_2, sum dE^2, 0 ; grad [B,3,H,W] or None)."""
    cam, ref_rgb, ref_lab = _f32c(cam), _f32c(ref_rgb), _f32c(ref_lab)
    B, _, H, W = cam.shape
    assert ref_rgb.shape == ref_lab.shape
    if stats is None:
        stats = torch.empty((B, 4), dtype=torch.float32, device=cam.device)
    if grad is None and want_grad:
        grad = torch.empty_like(cam)
    L = lib()
    ws = workspace("color_loss", L.spaa_color_loss_ws_bytes(B, H * W), cam.device)
    L.spaa_color_loss_fwd_bwd(_p(cam), _p(ref_rgb), _p(ref_lab), _bstride(ref_rgb, B), B, H * W, int(cam_is_lab2),
                              int(de_weighting), float(c_de), float(c_l2), int(bool(fast) and grad is not None), _p(stats), _p(grad), _p(ws), _stream()); _count()
    return stats, grad


# ------------------------------------------------------------------------------------------------------------
# warping
# ------------------------------------------------------------------------------------------------------------

def tps_grid(theta: Tensor, ctrl: Tensor, H: int, W: int) -> Tensor:
    """pytorch_tps.tps_grid as a planar [2,H,W] grid (x plane, y plane)."""
    theta, ctrl = _f32c(theta), _f32c(ctrl)
    T = ctrl.shape[0]
    assert theta.numel() == (T + 2) * 2, "only the reduced TPS form (T+2 rows) is implemented, as used by WarpingNet"
    g = torch.empty((2, H, W), dtype=torch.float32, device=theta.device)
    lib().spaa_tps_grid_fwd(_p(theta), _p(ctrl), T, H, W, _p(g), _stream()); _count()
    return g


def coarse_grid(affine: Tensor, theta: Tensor, ctrl: Tensor, in_hw, out_hw) -> Tensor:
    affine, theta, ctrl = _f32c(affine), _f32c(theta), _f32c(ctrl)
    T = ctrl.shape[0]
    g = torch.empty((2, out_hw[0], out_hw[1]), dtype=torch.float32, device=theta.device)
    lib().spaa_coarse_grid_fwd(_p(affine), _p(theta), _p(ctrl), T, in_hw[0], in_hw[1], out_hw[0], out_hw[1], _p(g), _stream()); _count()
    return g


def coarse_grid_bwd(affine: Optional[Tensor], theta: Tensor, ctrl: Tensor, in_hw, out_hw, dgrid: Tensor):
    theta, ctrl, dgrid = _f32c(theta), _f32c(ctrl), _f32c(dgrid)
    T = ctrl.shape[0]
    L = lib()
    daff = torch.empty((1, 2, 3), dtype=torch.float32, device=theta.device) if affine is not None else None
    dtheta = torch.empty_like(theta)
    ws = workspace("coarse_grid", L.spaa_coarse_grid_ws_bytes(T, out_hw[0], out_hw[1]), theta.device)
    L.spaa_coarse_grid_bwd(_p(_f32c(affine)) if affine is not None else None, _p(theta), _p(ctrl), T, in_hw[0], in_hw[1], out_hw[0],
                           out_hw[1], _p(dgrid), _p(daff), _p(dtheta), _p(ws), _stream()); _count()
    return daff, dtheta


def grid_finish(coarse: Tensor, refine: Optional[Tensor]) -> Tensor:
    coarse = _f32c(coarse)
    refine = _f32c(refine) if refine is not None else None
    fine = torch.empty_like(coarse)
    lib().spaa_grid_finish_fwd(_p(coarse), _p(refine), _p(fine), coarse.numel(), _stream()); _count()
    return fine


def grid_finish_bwd(coarse: Tensor, refine: Optional[Tensor], dfine: Tensor) -> Tensor:
    coarse, dfine = _f32c(coarse), _f32c(dfine)
    refine = _f32c(refine) if refine is not None else None
    d = torch.empty_like(coarse)
    lib().spaa_grid_finish_bwd(_p(coarse), _p(refine), _p(dfine), _p(d), coarse.numel(), _stream()); _count()
    return d


def _grid_bs(grid: Tensor, B: int) -> int:
    if grid.dim() == 3:
        return 0
    return _bstride(grid, B)


def _check_warp_shapes(grid: Tensor, B: int, H: int, W: int, mask, out, out_shape, rough, what: str = "grid_sample") -> None:
    """Where the reference would raise a shape error (F.grid_sample / broadcasting, models.py:184,340-342) the raw-pointer kernels would read
    or write out of bounds: check every operand against the grid's H x W."""
    if grid.dim() not in (3, 4) or grid.shape[-3] != 2 or (grid.dim() == 4 and grid.shape[0] not in (1, B)):
        raise ValueError(f"{what}: grid must be planar [2,H,W] or [B,2,H,W] (B = {B}); got {tuple(grid.shape)}")
    if mask is not None and mask.numel() != H * W:
        raise ValueError(f"{what}: mask has {mask.numel()} elements, the sampling grid is {H}x{W}")
    if tuple(out.shape) != tuple(out_shape):
        raise ValueError(f"{what}: output buffer has shape {tuple(out.shape)}, expected {tuple(out_shape)}")
    if rough is not None and (rough.dim() != 4 or rough.shape[0] not in (1, B) or tuple(rough.shape[-2:]) != (H, W)):
        raise ValueError(f"{what}: the surface image has shape {tuple(rough.shape)}, the sampling grid is {H}x{W} (batch {B})")


def grid_sample(img: Tensor, grid: Tensor, *, clamp01: bool = False, mask: Optional[Tensor] = None,
                out: Optional[Tensor] = None, rough: Optional[Tensor] = None, out2: Optional[Tensor] = None) -> Tensor:
    """img [B,C,Hi,Wi]; grid planar [2,H,W] (shared) or [B,2,H,W]; mask [H*W] floats; out [B,C,H,W].
    rough/out2: out2 = out * rough written in the same pass (out2 may be a channel slice of a wider tensor)."""
    img, grid = _f32c(img), _f32c(grid)
    B, C, Hi, Wi = img.shape
    H, W = grid.shape[-2:]
    if out is None:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=img.device)
    _check_warp_shapes(grid, B, H, W, mask, out, (B, C, H, W), rough if out2 is not None else None)
    assert out.is_contiguous()
    rb = ob = 0
    if out2 is not None:
        assert rough is not None and rough.dtype == torch.float32 and out2.dtype == torch.float32
        if tuple(out2.shape) != (B, C, H, W):
            raise ValueError(f"grid_sample: out2 has shape {tuple(out2.shape)}, expected {(B, C, H, W)}")
        assert rough.stride(1) == H * W and rough.stride(3) == 1 and out2.stride(1) == H * W and out2.stride(3) == 1
        rb, ob = _bstride(rough, B), out2.stride(0)
    lib().spaa_grid_sample_fwd(_p(img), B, C, Hi, Wi, _p(grid), _grid_bs(grid, B), H, W, int(clamp01), _p(mask), _p(out),
                               _p(rough) if out2 is not None else None, rb, _p(out2), ob, _stream()); _count()
    return out


def grid_sample_packed(img: Tensor, grid: Tensor, dtype, *, clamp01: bool = False, mask: Optional[Tensor] = None,
                       rough: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """3-channel warp emitted as the zero-padded 16-channel NHWC tensor [x | s | x*s | 0] (logical shape [B,16,H,W],
    channels-last) that the tensor-core conv1 / conv1_s read."""
    img, grid = _f32c(img), _f32c(grid)
    B, C, Hi, Wi = img.shape
    assert C == 3
    H, W = grid.shape[-2:]
    if out is None:
        out = torch.empty((B, 16, H, W), dtype=dtype, device=img.device, memory_format=torch.channels_last)
    _check_warp_shapes(grid, B, H, W, mask, out, (B, 16, H, W), rough)
    if not out.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("grid_sample_packed: `out` must be a dense channels_last [B,16,H,W] tensor")
    lib().spaa_grid_sample_fwd_packed(_p(img), B, Hi, Wi, _p(grid), _grid_bs(grid, B), H, W, int(clamp01), _p(mask), _p(rough),
                                      _bstride(rough, B) if rough is not None else 0, _p(out), _dt(out), _stream()); _count()
    return out


def pack_nhwc16(x: Tensor, surf: Optional[Tensor], dtype, split: bool = False, out: Optional[Tensor] = None) -> Tensor:
    """[x | surf | 0] as a zero-padded 16-channel NHWC tensor of `dtype` (logical shape [B,16,H,W], channels-last); split: the bf16x3
    split-precision form [h(16) | m(16) | l(16)] (logical shape [B,48,H,W])."""
    x = _f32c(x)
    B, Cx, H, W = x.shape
    Cs, sb = 0, 0
    if surf is not None:
        surf = _f32c(surf)
        Cs, sb = surf.shape[1], _bstride(surf, B)
    if split:
        if out is None:
            out = torch.empty((B, 48, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        assert out.shape == (B, 48, H, W) and out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=torch.channels_last)
        lib().spaa_pack_nhwc16_split3(_p(x), Cx, _p(surf), Cs, sb, _p(out), B, H * W, _stream()); _count()
        return out
    out = torch.empty((B, 16, H, W), dtype=dtype, device=x.device, memory_format=torch.channels_last)
    lib().spaa_pack_nhwc16(_p(x), Cx, _p(surf), Cs, sb, _p(out), _dt(out), B, H * W, _stream()); _count()
    return out


def half_to_bf16(x: Tensor) -> Tensor:
    """bf16 copy of a dense fp16 tensor (same memory format)."""
    assert x.dtype == torch.float16 and x.numel() % 8 == 0
    out = torch.empty_like(x, dtype=torch.bfloat16)           # preserve_format: same strides
    assert out.stride() == x.stride()
    lib().spaa_half_to_bf16(_p(x), _p(out), x.numel(), _stream()); _count()
    return out


def select_cotangent_packed(g0: Tensor, g1: Optional[Tensor], sel: Optional[Tensor], act: Optional[Tensor], mask_mode: int, out: Tensor) -> Tensor:
    """select_cotangent for 3-channel images, written as zero-padded 16-channel NHWC (`out`: [B,16,H,W] channels-last)."""
    B, C, H, W = g0.shape
    assert C == 3 and out.shape == (B, 16, H, W)
    lib().spaa_select_cotangent_packed(_p(g0), _p(g1), _p(sel), _p(act), mask_mode, _p(out), _dt(out), B, H * W, _stream()); _count()
    return out


def grid_sample_bwd_input(dout: Tensor, grid: Tensor, in_hw, *, mask: Optional[Tensor] = None, dout2: Optional[Tensor] = None,
                          rough: Optional[Tensor] = None, dimg: Optional[Tensor] = None) -> Tensor:
    dout, grid = _f32c(dout), _f32c(grid)
    B, C, H, W = dout.shape
    if dimg is None:
        dimg = torch.zeros((B, C, in_hw[0], in_hw[1]), dtype=torch.float32, device=dout.device)
    else:
        dimg.zero_()
    _check_warp_shapes(grid, B, H, W, mask, dimg, (B, C, int(in_hw[0]), int(in_hw[1])), rough if dout2 is not None else None, what="grid_sample_bwd_input")
    if tuple(grid.shape[-2:]) != (H, W):
        raise ValueError(f"grid_sample_bwd_input: the grid is {tuple(grid.shape[-2:])} but the output gradient is {(H, W)}")
    if dout2 is not None and tuple(dout2.shape) != (B, C, H, W):
        raise ValueError(f"grid_sample_bwd_input: dout2 has shape {tuple(dout2.shape)}, expected {(B, C, H, W)}")
    db = rb = 0
    if dout2 is not None:
        assert dout2.stride(1) == H * W and dout2.stride(3) == 1 and rough.stride(1) == H * W
        db, rb = dout2.stride(0), _bstride(rough, B)
    lib().spaa_grid_sample_bwd_input(_p(dout), _p(dout2), db, _p(rough) if dout2 is not None else None, rb, _p(mask), _p(grid),
                                     _grid_bs(grid, B), B, C, in_hw[0], in_hw[1], H, W, _p(dimg), _stream()); _count()
    return dimg


GATHER_TILED = os.environ.get("SPAA_GATHER_TILED", "1") not in ("", "0")      # tiled (shared-memory staged) form of the deterministic warp adjoint


class WarpAdjoint:
    """CSR form of the adjoint of a FIXED bilinear warp (planar grid [2,H,W] shared by the batch, optional mask [H*W]): built once per attack,
    then every backward is a gather (grid_sample_bwd_gather) -- deterministic, no atomics, no zero-fill, squared norm fused."""

    def __init__(self, grid: Tensor, in_hw, mask: Optional[Tensor] = None):
        grid = _f32c(grid)
        assert grid.dim() == 3 and grid.shape[0] == 2, "WarpAdjoint needs one planar grid [2,H,W] shared by the batch"
        _, H, W = grid.shape
        Hi, Wi = int(in_hw[0]), int(in_hw[1])
        HW, HWi = H * W, Hi * Wi
        dev = grid.device
        ent_q = torch.empty(4 * HW, dtype=torch.int32, device=dev)
        ent_w = torch.empty(4 * HW, dtype=torch.float32, device=dev)
        m = _f32c(mask).reshape(-1) if mask is not None else None
        lib().spaa_warp_taps(_p(grid), _p(m), Hi, Wi, H, W, _p(ent_q), _p(ent_w), _stream()); _count()
        # once per attack: stable sort by input pixel (fixes the summation order), CSR offsets by binary search -- plumbing, not the hot path
        q_sorted, order = torch.sort(ent_q.long(), stable=True)
        n_used = int((q_sorted < HWi).sum())
        order = order[:n_used]
        self.ent_p = (order % HW).to(torch.int32).contiguous()
        self.ent_w = ent_w[order].contiguous()
        self.ent_m = (m[self.ent_p.long()] if m is not None else torch.ones(n_used, device=dev)).contiguous()
        self.row_ptr = torch.searchsorted(q_sorted[:n_used].contiguous(), torch.arange(HWi + 1, device=dev)).to(torch.int32).contiguous()
        self.in_hw, self.out_hw = (Hi, Wi), (H, W)
        # tiled gather (spaa_grid_sample_bwd_gather_tiled): per 32 x 32 tile of input pixels the rectangle of output pixels that contribute to it,
        # and every entry's index inside its tile's rectangle
        GT = 32
        tx_n, ty_n = (Wi + GT - 1) // GT, (Hi + GT - 1) // GT
        q_used = q_sorted[:n_used]
        tile = (q_used // Wi) // GT * tx_n + (q_used % Wi) // GT
        py, px = self.ent_p.long() // W, self.ent_p.long() % W
        nt = tx_n * ty_n
        big = torch.full((nt,), 1 << 30, dtype=torch.long, device=dev)
        y0 = big.clone().scatter_reduce_(0, tile, py, "amin")
        x0 = big.clone().scatter_reduce_(0, tile, px, "amin")
        y1 = torch.full((nt,), -1, dtype=torch.long, device=dev).scatter_reduce_(0, tile, py, "amax")
        x1 = torch.full((nt,), -1, dtype=torch.long, device=dev).scatter_reduce_(0, tile, px, "amax")
        empty = y1 < 0
        hh = torch.where(empty, torch.zeros_like(y1), y1 - y0 + 1)
        ww = torch.where(empty, torch.zeros_like(x1), x1 - x0 + 1)
        y0, x0 = torch.where(empty, torch.zeros_like(y0), y0), torch.where(empty, torch.zeros_like(x0), x0)
        self.boxes = torch.stack((y0, x0, hh, ww), 1).to(torch.int32).contiguous()
        self.ent_l = ((py - y0[tile]) * ww[tile] + (px - x0[tile])).to(torch.int32).contiguous()
        self.max_region = int((hh * ww).max().item()) if nt else 0


def grid_sample_bwd_gather(adj: WarpAdjoint, dout: Tensor, *, dout2: Optional[Tensor] = None, rough: Optional[Tensor] = None,
                           dimg: Optional[Tensor] = None, sq: Optional[Tensor] = None, x_for_clamp: Optional[Tensor] = None, lo: float = 0.0,
                           hi: float = 1.0) -> Tensor:
    """dimg = adjoint-warp of (dout + dout2 * rough) * mask through `adj`; sq[b] (optional) = squared norm of dimg[b] over the entries whose
    x_for_clamp lies in [lo,hi] (all entries without it)."""
    dout = _f32c(dout)
    B, C, H, W = dout.shape
    assert (H, W) == adj.out_hw and C <= 3
    Hi, Wi = adj.in_hw
    if dimg is None:
        dimg = torch.empty((B, C, Hi, Wi), dtype=torch.float32, device=dout.device)
    db = rb = 0
    if dout2 is not None:
        assert dout2.stride(1) == H * W and dout2.stride(3) == 1 and rough.stride(1) == H * W
        db, rb = dout2.stride(0), _bstride(rough, B)
    L = lib()
    if GATHER_TILED and 0 < adj.max_region * C * 4 <= 160 * 1024:
        ws = workspace("gs_gather_tiled", L.spaa_grid_sample_bwd_gather_tiled_ws_bytes(B, Hi, Wi), dout.device) if sq is not None else None
        L.spaa_grid_sample_bwd_gather_tiled(_p(dout), _p(dout2), db, _p(rough) if dout2 is not None else None, rb, _p(adj.row_ptr), _p(adj.ent_l), _p(adj.ent_w),
                                            _p(adj.ent_m), _p(adj.boxes), adj.max_region, B, C, Hi, Wi, H, W,
                                            _p(_f32c(x_for_clamp)) if x_for_clamp is not None else None, float(lo), float(hi), _p(dimg), _p(sq), _p(ws), _stream()); _count()
        return dimg
    ws = workspace("gs_gather", L.spaa_grid_sample_bwd_gather_ws_bytes(B, Hi, Wi), dout.device) if sq is not None else None
    L.spaa_grid_sample_bwd_gather(_p(dout), _p(dout2), db, _p(rough) if dout2 is not None else None, rb, _p(adj.row_ptr), _p(adj.ent_p), _p(adj.ent_w),
                                  _p(adj.ent_m), B, C, Hi, Wi, H, W, _p(_f32c(x_for_clamp)) if x_for_clamp is not None else None, float(lo), float(hi),
                                  _p(dimg), _p(sq), _p(ws), _stream()); _count()
    return dimg


def grid_sample_bwd_grid(dout: Tensor, img: Tensor, grid: Tensor, *, clamp01: bool = False, mask: Optional[Tensor] = None,
                         dout2: Optional[Tensor] = None, rough: Optional[Tensor] = None) -> Tensor:
    dout, img, grid = _f32c(dout), _f32c(img), _f32c(grid)
    B, C, H, W = dout.shape
    Hi, Wi = img.shape[-2:]
    shared = grid.dim() == 3 or grid.shape[0] == 1
    dgrid = torch.empty((2, H, W) if shared else (B, 2, H, W), dtype=torch.float32, device=dout.device)
    db = rb = 0
    if dout2 is not None:
        db, rb = dout2.stride(0), _bstride(rough, B)
    lib().spaa_grid_sample_bwd_grid(_p(dout), _p(dout2), db, _p(rough) if dout2 is not None else None, rb, _p(mask), _p(img),
                                    int(clamp01), _p(grid), 0 if shared else grid.stride(0), B, C, Hi, Wi, H, W, _p(dgrid),
                                    _stream()); _count()
    return dgrid


# ------------------------------------------------------------------------------------------------------------
# convolution (gather conv, see include/spaa_b200.h)
# ------------------------------------------------------------------------------------------------------------

class ConvSpec:
    """Static description of an nn.Conv2d (kind 'conv') or nn.ConvTranspose2d (kind 'convT') layer."""

    def __init__(self, kind: str, cin: int, cout: int, k: int, stride: int = 1, pad: int = 0, outpad: int = 0):
        assert kind in ("conv", "convT")
        self.kind, self.cin, self.cout, self.k, self.stride, self.pad, self.outpad = kind, cin, cout, k, stride, pad, outpad

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        if self.kind == "conv":
            f = lambda n: (n + 2 * self.pad - self.k) // self.stride + 1
        else:
            f = lambda n: (n - 1) * self.stride - 2 * self.pad + self.k + self.outpad
        return f(h), f(w)

    def weight_shape(self):
        return (self.cout, self.cin, self.k, self.k) if self.kind == "conv" else (self.cin, self.cout, self.k, self.k)


_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _dt(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported activation dtype {t.dtype}") from None


def _act_strides(t: Tensor) -> Tuple[int, int, int]:
    """(batch, pixel, channel) element strides of a logical [B,C,H,W] tensor (NCHW or channels-last, or a channel slice)."""
    sb, sc, sh, sw = t.stride()
    if t.shape[2] > 1 and sh != t.shape[3] * sw:
        raise ValueError("activation rows must be densely packed (stride_h == W * stride_w)")
    return (0 if t.shape[0] == 1 else sb), sw, sc


def _fill_desc(d: ConvDesc, x: Tensor, out: Tensor, add: Optional[Tensor], mask: Optional[Tensor], split: bool = False) -> None:
    d.split = int(bool(split))
    d.in_dtype, d.out_dtype = _dt(x), _dt(out)
    d.B = out.shape[0]
    d.Hin, d.Win = x.shape[2], x.shape[3]
    d.Hout, d.Wout = out.shape[2], out.shape[3]
    d.in_bs, d.in_ps, d.in_cs = _act_strides(x)
    d.out_bs, d.out_ps, d.out_cs = _act_strides(out)
    if add is not None:
        assert add.dtype == out.dtype and add.shape[1:] == out.shape[1:]
        d.add_bs, d.add_ps, d.add_cs = _act_strides(add)
    if mask is not None:
        assert mask.shape[1:] == out.shape[1:]      # dtype may differ between 16-bit types (tensor-core path: sign test only)
        d.mask_bs, d.mask_ps, d.mask_cs = _act_strides(mask)


def _w_strides(w: Tensor, k: int):
    assert w.dtype == torch.float32 and w.stride(3) == 1 and (k == 1 or w.stride(2) == k), "weights must be fp32 with dense taps"
    return w.stride(0), w.stride(1)


_graph_pools: Dict[int, object] = {}


def graph_pool(device) -> Optional[object]:
    """OPT-IN ($SPAA_GRAPH_POOL=1; default: every graph has its own pool): one CUDA-graph memory pool per device, shared by the captures of the attack
    engines (`CUDAGraph.capture_begin(pool=...)`).  Measured both ways on B200: the FIRST spaa() call of a process took 0.66-0.67 s with it and
    1.1-3.6 s without (a capture with a pool of its own takes its ~1.4 GB of per-iteration activations from fresh cudaMalloc segments), but the blocks of
    a destroyed graph are not handed to the next capture either way (tools/capture_probe.py: +1.4 GB reserved per engine), so with the pool kept alive
    they pile up until release_graph_pools(), and PerC-AL -- one short-lived graph per call -- ran at HALF speed with it (inception_v3: 26.7 vs 54.3
    it/s).  Hence off by default.  The usual rule for shared pools applies: graphs are replayed on one stream, never concurrently, and a tensor PRODUCED
    by a replay is read before another engine replays."""
    if os.environ.get("SPAA_GRAPH_POOL", "0") == "0":
        return None
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _graph_pools:
        # the allocator drops a pool when the last graph captured in it is destroyed, and a stale handle then trips an internal assert on the next
        # capture: a tiny anchor graph captured in the pool keeps it alive until release_graph_pools()
        handle = torch.cuda.graph_pool_handle()
        anchor = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=idx)
        side.wait_stream(torch.cuda.current_stream(idx))
        with torch.cuda.device(idx), torch.cuda.stream(side):
            anchor.capture_begin(pool=handle)
            try:
                keep = torch.zeros(8, device=f"cuda:{idx}")
            finally:
                anchor.capture_end()
        torch.cuda.current_stream(idx).wait_stream(side)
        _graph_pools[idx] = (handle, anchor, keep, side)
    return _graph_pools[idx][0]


def release_graph_pools() -> None:
    """Drop the shared graph pools (after the engines that captured into them are gone): their memory returns to the caching allocator."""
    _graph_pools.clear()


_packed_cache: Dict[tuple, Tensor] = {}
TC_ENABLED = True          # tests flip this to compare the tensor-core path against the CUDA-core path
WGRAD_SCRATCH = os.environ.get("SPAA_WGRAD_SCRATCH", "1") != "0"      # backward-weight flush through the scratch gradient (vector reductions)


_weights_epoch = 0


def weights_epoch() -> int:
    """Incremented whenever parameters were updated behind autograd's back (raw Adam kernel): part of the attack engines' cache key."""
    return _weights_epoch


def invalidate_packed_weights() -> None:
    """Drop the 16-bit packed-weight cache and bump the weights epoch (call after parameters were updated in place by a raw kernel, which
    leaves their autograd version counters unchanged): engines / CUDA graphs built from the previous weights are then never reused."""
    global _weights_epoch, _repack_table
    _weights_epoch += 1
    _packed_cache.clear()
    _repack_table = None


_repack_table = None          # (cache keys it covers, device table of pack jobs, njobs)
_repack_keep: list = []


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


def repack_packed_weights() -> None:
    """After an in-place parameter update by a raw kernel (FlatAdam): refresh EVERY cached 16-bit weight copy from its fp32 parameter with ONE
    launch (spaa_conv_tc_pack_weights_multi) instead of dropping the cache and re-packing layer by layer (34 launches per training step).  The
    packed tensors keep their addresses, so a CUDA graph that recorded this call stays valid.  Bumps the weights epoch like
    invalidate_packed_weights()."""
    global _weights_epoch, _repack_table
    _weights_epoch += 1
    for k in [k for k, v in _packed_cache.items() if v[3] is None or v[0]() is None]:      # split-precision copies / dead parameters: rebuilt lazily
        del _packed_cache[k]
    if not _packed_cache:
        _repack_table = None
        return
    if _repack_table is None or _repack_table[0] != tuple(_packed_cache.keys()):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the set of packed weights changed inside a CUDA-graph capture")
        dev = next(iter(_packed_cache.values()))[2].device
        blob = b"".join(v[3] for v in _packed_cache.values())
        table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        _repack_table = (tuple(_packed_cache.keys()), table, len(_packed_cache))
        _repack_keep.append((table, [v[2] for v in _packed_cache.values()]))       # a captured CUDA graph may hold these addresses
        del _repack_keep[:-8]
    _, table, n = _repack_table
    lib().spaa_conv_tc_pack_weights_multi(_p(table), n, _stream()); _count()


def _tc_weights(d: ConvDesc, w: Tensor, cin_real: int, cin_off: int) -> Tensor:
    """Packed 16-bit copy of a layer's weights for the tensor-core kernel, cached per parameter tensor.  The entry holds a
    weak reference to the parameter (the base of `w` when `w` is a channel slice): a different tensor that happens to be
    allocated at the same address after the first one was freed can never hit a stale entry, and in-place updates are
    caught by the version counter."""
    base = w._base if w._base is not None else w
    key = (id(base), w.data_ptr(), tuple(w.shape), tuple(w.stride()), d.flip, d.w_cis, d.w_cos, d.KH, d.Cin, d.Cout, d.in_dtype, cin_real, cin_off, d.split)
    hit = _packed_cache.get(key)
    if hit is not None and hit[0]() is base and hit[1] == base._version:
        return hit[2]
    L = lib()
    t = torch.empty(L.spaa_conv_tc_packed_elems(ctypes.byref(d)), dtype=torch.int16, device=w.device)
    job = None
    if d.split:
        w6, d6 = _split_weights(d, w, cin_real, cin_off)
        L.spaa_conv_tc_pack_weights(ctypes.byref(d6), _p(w6), 6 * d.Cin, 0, _p(t), _stream()); _count()
    else:
        L.spaa_conv_tc_pack_weights(ctypes.byref(d), _p(w), cin_real, cin_off, _p(t), _stream()); _count()
        buf = ctypes.create_string_buffer(int(L.spaa_conv_tc_pack_job_bytes()))
        L.spaa_conv_tc_pack_job(ctypes.byref(d), _p(w), cin_real, cin_off, _p(t), buf)
        job = buf.raw
    if len(_packed_cache) > 512:
        for k in [k for k, v in _packed_cache.items() if v[0]() is None]:
            del _packed_cache[k]
        if len(_packed_cache) > 512:
            _packed_cache.clear()
    _packed_cache[key] = (weakref.ref(base), base._version, t, job)
    return t


# bf16x3 split-precision mode (spaa_conv_desc.split): the six part products of input x filter in the order the kernel's chunk table walks the INPUT
# parts (conv_tc.cu: a_part = l, h, m, m, h, h): smallest products first
_SPLIT_W_PART = (0, 2, 1, 0, 1, 0)          # filter part (0 = h, 1 = m, 2 = l) of product k


def split3(v: Tensor):
    """(h, m, l): the three bf16 parts of an fp32 tensor, as fp32 tensors; h + m + l == v to 24 significand bits."""
    h = v.bfloat16().float()
    m = (v - h).bfloat16().float()
    return h, m, (v - h - m).bfloat16().float()


def _split_weights(d: ConvDesc, w: Tensor, cin_real: int, cin_off: int):
    """fp32 filter -> the 6 * Cin input-channel filter of the split mode (each block: one bf16 part of the filter, the real channels at
    [cin_off, cin_off + cin_real) of the zero-padded Cin) + the descriptor that packs it with the ordinary pack kernel."""
    ci_dim = 0 if d.w_cis == w.stride(0) else 1
    if w.shape[ci_dim] != cin_real:
        raise RuntimeError("split-precision weights: cannot identify the contracted dimension of the filter")
    parts = split3(w.detach().float())
    shape = list(w.shape)
    shape[ci_dim] = 6 * d.Cin
    w6 = torch.zeros(shape, dtype=torch.float32, device=w.device)
    for k, pb in enumerate(_SPLIT_W_PART):
        w6.narrow(ci_dim, k * d.Cin + cin_off, cin_real).copy_(parts[pb])
    d6 = ConvDesc()
    ctypes.memmove(ctypes.byref(d6), ctypes.byref(d), ctypes.sizeof(ConvDesc))
    d6.split, d6.Cin = 0, 6 * d.Cin
    d6.w_ts = 1
    if ci_dim == 0:
        d6.w_cis, d6.w_cos = w6.stride(0), w6.stride(1)
    else:
        d6.w_cis, d6.w_cos = w6.stride(1), w6.stride(0)
    return w6, d6


def _launch_conv(kind: str, spec, d: ConvDesc, x, w, b, add, mask, mask2, out, out2, cin_real: Optional[int] = None, cin_off: int = 0) -> None:
    L = lib()
    planar = d.out_dtype == 0
    allowed = EPI_RELU | (EPI_CLAMP_MAX1 if planar else EPI_OUT2_BF16)
    use_tc = (TC_ENABLED and d.in_dtype in (1, 2) and (d.epi_flags & ~allowed) == 0 and (mask is None or d.mask_mode == MASK_POS)
              and not (planar and (mask is not None or mask2 is not None)) and L.spaa_conv_tc_supported(ctypes.byref(d)) == 1)
    padded = cin_real is not None and cin_real != d.Cin
    if not use_tc:
        if d.split:
            raise RuntimeError("split-precision (bf16x3) operands exist on the tensor-core path only, and this layer shape is not covered by it")
        if padded:
            raise RuntimeError("zero-padded channel inputs are only implemented on the tensor-core path")
        for m in (mask, mask2):
            if m is not None and m.dtype != out.dtype:
                raise RuntimeError("the CUDA-core conv path needs masks of the output's dtype")
    with _Probe(kind + ("_tc" if use_tc else ""), spec):
        if use_tc:
            wp = _tc_weights(d, w, cin_real if cin_real is not None else d.Cin, cin_off)
            L.spaa_conv_tc_fwd(ctypes.byref(d), _p(x), _p(wp), _p(b), _p(add), _p(mask), _p(mask2), _p(out), _p(out2), _stream())
        else:
            L.spaa_conv_fwd(ctypes.byref(d), _p(x), _p(w), _p(b), _p(add), _p(mask), _p(mask2), _p(out), _p(out2), _stream())
    _count()


def _new_act(shape, dtype, device) -> Tensor:
    """Activation buffer: bf16 tensors are NHWC (channels-last) so the tensor-core kernels can TMA them; fp32 stays NCHW."""
    if dtype in (torch.bfloat16, torch.float16):
        return torch.empty(shape, dtype=dtype, device=device, memory_format=torch.channels_last)
    return torch.empty(shape, dtype=dtype, device=device)


def conv_forward(spec: ConvSpec, x: Tensor, w: Tensor, b: Optional[Tensor], *, out: Optional[Tensor] = None,
                 add: Optional[Tensor] = None, epi: int = 0, out_dtype=None, cin_offset: int = 0, split: bool = False,
                 bf16_copy: Optional[dict] = None) -> Tensor:
    """Forward of nn.Conv2d / nn.ConvTranspose2d with the fused epilogue `epi` (bias, residual add, activation, clamp).
    bf16_copy: a dict that receives {out.data_ptr(): bf16 rounding of the same result} when the output is fp16 NHWC (the fp16 training mode:
    the backward-weight kernel wants its activation operand in the gradients' format); written by the same epilogue on the tensor-core path.
    `x` may carry more channels than spec.cin (a zero-padded 16-channel NHWC tensor): the layer then reads channels
    [cin_offset, cin_offset + spec.cin) (tensor-core path only).
    split: bf16x3 split-precision operands (tensor-core path): x / add / a 16-bit out carry three bf16 parts per logical channel
    ([h | m | l], 3x the channels); an fp32 `out_dtype` output is the plain fp32 result."""
    _need_cuda(x, w)
    B, _, H, W = x.shape
    Ho, Wo = spec.out_hw(H, W)
    np_ = 3 if split else 1
    odt = out_dtype or x.dtype
    if out is None:
        out = _new_act((B, spec.cout * (np_ if odt != torch.float32 else 1), Ho, Wo), odt, x.device)
    d = ConvDesc()
    d.Cin, d.Cout, d.KH, d.KW = x.shape[1] // np_, spec.cout, spec.k, spec.k
    s0, s1 = _w_strides(w, spec.k)
    d.w_ts = 1
    if spec.kind == "conv":
        d.stride, d.up, d.pad_h, d.pad_w, d.flip = spec.stride, 1, spec.pad, spec.pad, 0
        d.w_cos, d.w_cis = s0, s1
    else:
        d.stride, d.up, d.flip = 1, spec.stride, 1
        d.pad_h = d.pad_w = spec.k - 1 - spec.pad
        d.w_cis, d.w_cos = s0, s1
    d.epi_flags, d.mask_mode = epi, MASK_NONE
    _fill_desc(d, x, out, add, None, split)
    out2 = None
    if bf16_copy is not None and out.dtype == torch.float16 and not split:
        if TC_ENABLED and x.dtype == torch.float16 and lib().spaa_conv_tc_supported(ctypes.byref(d)) == 1:
            out2 = torch.empty_like(out, dtype=torch.bfloat16)
            assert out2.stride() == out.stride()
            d.epi_flags = epi | EPI_OUT2_BF16
    _launch_conv("fwd", spec, d, x, w, b, add, None, None, out, out2, cin_real=spec.cin, cin_off=cin_offset)
    if bf16_copy is not None and out.dtype == torch.float16 and not split:
        bf16_copy[out.data_ptr()] = out2 if out2 is not None else half_to_bf16(out)
    return out


def conv_backward_data(spec: ConvSpec, dy: Tensor, w: Tensor, in_hw, *, out: Optional[Tensor] = None, add: Optional[Tensor] = None,
                       mask: Optional[Tensor] = None, mask_mode: int = MASK_NONE, mask2: Optional[Tensor] = None,
                       out2: Optional[Tensor] = None, out_dtype=None, split: bool = False) -> Tensor:
    """Gradient wrt the layer input: out = mask(mask_mode) * (bwd_data(dy) + add); out2 = out * (mask2 > 0).
    `w` may be a channel-sliced view of the parameter (to produce only some input channels' gradients)."""
    _need_cuda(dy, w)
    B = dy.shape[0]
    cin = w.shape[1] if spec.kind == "conv" else w.shape[0]      # possibly sliced
    np_ = 3 if split else 1
    odt = out_dtype or dy.dtype
    if out is None:
        out = _new_act((B, cin * (np_ if odt != torch.float32 else 1), in_hw[0], in_hw[1]), odt, dy.device)
    d = ConvDesc()
    d.Cin, d.Cout, d.KH, d.KW = dy.shape[1] // np_, cin, spec.k, spec.k      # dy may be zero-padded to 16 channels (tensor-core path)
    s0, s1 = _w_strides(w, spec.k)
    d.w_ts = 1
    if spec.kind == "conv":       # dX = gather conv of dY with flipped taps, up = stride
        d.stride, d.up, d.flip = 1, spec.stride, 1
        d.pad_h = d.pad_w = spec.k - 1 - spec.pad
        d.w_cis, d.w_cos = s0, s1
    else:                         # ConvTranspose backward-data is a plain strided conv
        d.stride, d.up, d.pad_h, d.pad_w, d.flip = spec.stride, 1, spec.pad, spec.pad, 0
        d.w_cos, d.w_cis = s0, s1
    d.epi_flags, d.mask_mode = 0, (mask_mode if mask is not None else MASK_NONE)
    _fill_desc(d, dy, out, add, mask if mask is not None else mask2, split)
    if mask2 is not None:
        assert out2 is not None and out2.stride() == out.stride()
        if mask is not None:
            assert mask2.stride()[1:] == mask.stride()[1:]
    _launch_conv("bwd_data", spec, d, dy, w, None, add, mask, mask2, out, out2, cin_real=spec.cout, cin_off=0)
    return out


def _dense_nhwc16(t: Tensor, np_: int = 1) -> bool:
    return t.dtype in (torch.bfloat16, torch.float16) and t.is_contiguous(memory_format=torch.channels_last) and t.shape[1] % np_ == 0 and \
        t.shape[1] // np_ in (16, 32, 64, 128, 256)


_SPLIT_WGRAD_PARTS = ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0))        # (x part, dy part) of the six part products, smallest first


class WgradScratch:
    """Scratch gradients of the tensor-core backward-weight kernel (spaa_conv_wgrad_tc_scratch: layout [tap][X channel][DY channel], so that the
    kernel's flush is 16-byte vector reductions) for the layers of one backward pass, and the ONE launch that adds them into the parameter gradients
    (spaa_wgrad_scatter_multi, which also zeroes the scratch again).  One persistent zero-filled buffer per device, bump-allocated per pass."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf: Optional[Tensor] = None
        self.old: list = []                 # outgrown buffers with pending jobs
        self.off = 0
        self.jobs: list = []

    def take(self, n: int) -> Tensor:
        n_al = (n + 63) // 64 * 64
        if self.buf is None or self.off + n_al > self.buf.numel():
            if self.buf is not None:
                self.old.append(self.buf)
            self.buf = torch.zeros(max(4 << 20, 2 * n_al, 2 * (self.buf.numel() if self.buf is not None else 0)), dtype=torch.float32, device=self.device)
            self.off = 0
        sl = self.buf[self.off:self.off + n]
        self.off += n_al
        return sl

    def add(self, scratch: Tensor, dw: Tensor, ntap: int, cx_real: int, cy: int, cy_real: int, w_ts: int, w_xs: int, w_ys: int) -> None:
        self.jobs.append((scratch, dw, ntap, cx_real, cy, cy_real, w_ts, w_xs, w_ys))

    def flush(self) -> None:
        jobs, self.jobs = self.jobs, []
        for i in range(0, len(jobs), 24):
            part = jobs[i:i + 24]
            n = len(part)
            arr = lambda ct, k: (ct * n)(*[j[k] for j in part])
            sc = (ctypes.c_void_p * n)(*[j[0].data_ptr() for j in part])
            dws = (ctypes.c_void_p * n)(*[j[1].data_ptr() for j in part])
            with torch.cuda.device(self.device):
                lib().spaa_wgrad_scatter_multi(sc, dws, arr(ctypes.c_int32, 2), arr(ctypes.c_int32, 3), arr(ctypes.c_int32, 4), arr(ctypes.c_int32, 5),
                                               arr(ctypes.c_int64, 6), arr(ctypes.c_int64, 7), arr(ctypes.c_int64, 8), n, _stream())
            _count()
        self.off = 0
        self.old.clear()


_wgrad_scratch: Dict[int, WgradScratch] = {}


def wgrad_scratch(device) -> WgradScratch:
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _wgrad_scratch:
        _wgrad_scratch[idx] = WgradScratch(torch.device("cuda", idx))
    return _wgrad_scratch[idx]


def conv_backward_weight(spec: ConvSpec, x: Tensor, dy: Tensor, dw: Tensor, db: Optional[Tensor], *, x_offset: int = 0, split: bool = False,
                         defer_bias: Optional[list] = None, scratch: Optional[WgradScratch] = None) -> None:
    """dw += d(loss)/d(weight), db += d(loss)/d(bias); dw/db are fp32 accumulators in the parameter's own layout.
    16-bit dense NHWC operands go to the tcgen05 backward-weight kernel; `x` / `dy` may then be zero-padded to 16 channels
    (the layer's real input channels sit at [x_offset, x_offset + spec.cin) of `x`, the real output channels at [0, spec.cout) of `dy`).
    split: bf16x3 split-precision operands (three bf16 parts per logical channel): the six leading part products are six launches of the same
    kernel on the parts, all accumulating into dw."""
    _need_cuda(x, dy, dw)
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    np_ = 3 if split else 1
    d = ConvDesc()
    d.KH = d.KW = spec.k
    d.stride, d.up, d.pad_h, d.pad_w, d.flip = spec.stride, 1, spec.pad, spec.pad, 0
    d.w_ts = 1
    L = lib()
    use_tc = TC_ENABLED and _dense_nhwc16(x, np_) and _dense_nhwc16(dy, np_) and x.dtype == dy.dtype      # tcgen05 kind::f16: both operands one format
    if split and not use_tc:
        raise RuntimeError("split-precision (bf16x3) operands exist on the tensor-core path only")
    cx_l, cy_l = x.shape[1] // np_, dy.shape[1] // np_                     # logical (per-part) channel counts
    if not use_tc and (x.shape[1] != spec.cin or dy.shape[1] != spec.cout):
        raise RuntimeError("zero-padded channel operands are only implemented on the tensor-core backward-weight path")
    if spec.kind == "conv":
        d.Cin, d.Cout = cx_l, cy_l
        d.w_cos, d.w_cis = dw.stride(0), dw.stride(1)
        _fill_desc(d, x, dy, None, None, split)
        gathered, pointwise, g_real, g_off, p_real = x, dy, spec.cin, x_offset, spec.cout
    else:
        # dWt[ci,co,r,s] = sum in[iy,ci] * dOut[iy*s-p+r, co]: the gathered operand is dOut, the pointwise one is `x`
        d.Cin, d.Cout = cy_l, cx_l
        d.w_cis, d.w_cos = dw.stride(1), dw.stride(0)
        _fill_desc(d, dy, x, None, None, split)
        gathered, pointwise, g_real, g_off, p_real = dy, x, spec.cout, 0, spec.cin
    if use_tc and L.spaa_conv_wgrad_tc_supported(ctypes.byref(d)) == 1:
        # the kernel's final flush: into the parameter gradient with scalar atomics, or -- the training loops pass `scratch` -- into a scratch gradient
        # with the DY channel contiguous (vector reductions), added to `dw` by scratch.flush() once per backward pass; WGRAD_SCRATCH = False: always direct
        own = None
        if scratch is None and WGRAD_SCRATCH:
            scratch = own = wgrad_scratch(dw.device)
        if not WGRAD_SCRATCH:
            scratch = None
        sc = None
        if scratch is not None:
            sc = scratch.take(spec.k * spec.k * g_real * d.Cout)
            scratch.add(sc, dw, spec.k * spec.k, g_real, d.Cout, p_real, d.w_ts, d.w_cis, d.w_cos)
        fn = L.spaa_conv_wgrad_tc if sc is None else L.spaa_conv_wgrad_tc_scratch
        dst = _p(dw) if sc is None else _p(sc)
        with _Probe("bwd_weight_tc", spec):
            if split:
                cg, cp = gathered.shape[1] // 3, pointwise.shape[1] // 3
                for pg_, pp_ in _SPLIT_WGRAD_PARTS:
                    fn(ctypes.byref(d), ctypes.c_void_p(gathered.data_ptr() + pg_ * cg * 2), ctypes.c_void_p(pointwise.data_ptr() + pp_ * cp * 2),
                       dst, g_real, g_off, p_real, _stream())
                    _count()
            else:
                fn(ctypes.byref(d), _p(gathered), _p(pointwise), dst, g_real, g_off, p_real, _stream())
                _count()
        if own is not None:
            own.flush()                      # stand-alone call: add it into dw right away
        if db is not None:
            if split:
                # bias gradient = sum over pixels of dy = of its three parts (real channels at [0, spec.cout) of each part)
                bs, ps, cs = _act_strides(dy)
                for part in range(3):
                    L.spaa_channel_sum(ctypes.c_void_p(dy.data_ptr() + part * cy_l * 2), _dt(dy), dy.shape[0], db.numel(), dy.shape[2] * dy.shape[3], bs, ps, cs,
                                       _p(db), _stream()); _count()
            elif defer_bias is not None and _dense_nhwc16(dy):
                defer_bias.append((dy, db))          # summed with the other layers' bias gradients in one launch: channel_sum_multi
            elif spec.kind == "conv":
                L.spaa_channel_sum_nhwc16(_p(dy), _dt(dy), dy.shape[0] * dy.shape[2] * dy.shape[3], dy.shape[1], _p(_bias_scratch(db, dy.shape[1])), _stream()); _count()
                _bias_fold(db, dy.shape[1])
            else:
                channel_sum(dy, db)
        return
    if split:
        raise RuntimeError("split-precision backward-weight: this layer shape is not covered by the tensor-core kernel")
    if x.shape[1] != spec.cin or dy.shape[1] != spec.cout:
        raise RuntimeError("zero-padded channel operands are only implemented on the tensor-core backward-weight path")
    if spec.kind == "conv":
        L.spaa_conv_bwd_weight(ctypes.byref(d), _p(x), _p(dy), _p(dw), _p(db), _stream()); _count(2 if db is not None else 1)
    else:
        L.spaa_conv_bwd_weight(ctypes.byref(d), _p(dy), _p(x), _p(dw), None, _stream()); _count()
        if db is not None:
            channel_sum(dy, db)


_bias_tmp: Dict[Tuple[int, int], Tensor] = {}


def _bias_scratch(db: Tensor, c: int) -> Tensor:
    """Bias gradients of zero-padded outputs (3 real of 16 channels) are summed into a padded scratch vector first."""
    if db.numel() == c:
        return db
    key = (db.device.index or 0, c)
    t = _bias_tmp.get(key)
    if t is None:
        t = torch.zeros(c, dtype=torch.float32, device=db.device)
        _bias_tmp[key] = t
    else:
        t.zero_()
    return t


def _bias_fold(db: Tensor, c: int) -> None:
    if db.numel() != c:
        db += _bias_tmp[(db.device.index or 0, c)][:db.numel()]


def channel_sum_multi(jobs: list) -> None:
    """[(x, out)]: out[c] += sum over batch and pixels of x[:, c] for c < out.numel(), every x a dense 16-bit NHWC tensor (possibly zero-padded
    beyond out.numel() channels); one launch per 24 tensors of one dtype (the bias gradients of a whole backward pass were 17 launches)."""
    for dt in (torch.bfloat16, torch.float16):
        sel = [(x, o) for x, o in jobs if x.dtype == dt]
        for i in range(0, len(sel), 24):
            part = sel[i:i + 24]
            n = len(part)
            xs = (ctypes.c_void_p * n)(*[x.data_ptr() for x, _ in part])
            outs = (ctypes.c_void_p * n)(*[o.data_ptr() for _, o in part])
            npix = (ctypes.c_int64 * n)(*[x.shape[0] * x.shape[2] * x.shape[3] for x, _ in part])
            C = (ctypes.c_int32 * n)(*[x.shape[1] for x, _ in part])
            cr = (ctypes.c_int32 * n)(*[o.numel() for _, o in part])
            for x, o in part:
                assert _dense_nhwc16(x) and o.dtype == torch.float32 and o.is_contiguous() and o.numel() <= x.shape[1] and o.device == x.device
            with torch.cuda.device(part[0][0].device):
                lib().spaa_channel_sum_nhwc16_multi(xs, npix, C, cr, outs, _dt(part[0][0]), n, _stream())
            _count()


def channel_sum(x: Tensor, out: Tensor) -> None:
    """out[c] += sum over batch and pixels of x[:, c]."""
    if _dense_nhwc16(x) and out.numel() == x.shape[1]:
        lib().spaa_channel_sum_nhwc16(_p(x), _dt(x), x.shape[0] * x.shape[2] * x.shape[3], x.shape[1], _p(out), _stream()); _count()
        return
    bs, ps, cs = _act_strides(x)
    lib().spaa_channel_sum(_p(x), _dt(x), x.shape[0], x.shape[1], x.shape[2] * x.shape[3], bs, ps, cs, _p(out), _stream()); _count()


# ------------------------------------------------------------------------------------------------------------
# training loss / optimiser
# ------------------------------------------------------------------------------------------------------------

def ssim_l1(pred: Tensor, target: Tensor, w_l1: float, w_l2: float, w_ssim: float, *, want_grad: bool = True,
            want_map: bool = False, cot_map: Optional[Tensor] = None):
    """Returns (sums[4] = sum|d|, sum d^2, sum ssim, 0 ; grad or None ; ssim_map or None)."""
    pred, target = _f32c(pred), _f32c(target)
    assert pred.shape == target.shape
    B, C, H, W = pred.shape
    L = lib()
    sums = torch.empty(4, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    smap = torch.empty_like(pred) if want_map else None
    ws = workspace("ssim", L.spaa_ssim_l1_ws_bytes(B * C, H, W), pred.device)
    L.spaa_ssim_l1_fwd_bwd(_p(pred), _p(target), B * C, H, W, float(w_l1), float(w_l2), float(w_ssim),
                           _p(_f32c(cot_map)) if cot_map is not None else None, _p(sums), _p(smap), _p(grad), _p(ws), _stream()); _count()
    return sums, grad, smap


def adam_step(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, seg_end: Tensor, seg_lr: Tensor, seg_wd: Tensor, step: int,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    for t in (param, grad, m, v):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    lib().spaa_adam_step(_p(param), _p(grad), _p(m), _p(v), param.numel(), _p(seg_end), _p(seg_lr), _p(seg_wd), seg_end.numel(),
                         beta1, beta2, eps, int(step), float(grad_scale), _stream()); _count()


def adam_step_dev(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, seg_end: Tensor, seg_lr: Tensor, seg_wd: Tensor, bias_corr2: Tensor,
                  beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    """adam_step with the per-step scalars (bias corrections, learning rates) in device buffers: CUDA-graph replayable."""
    for t in (param, grad, m, v, seg_lr, bias_corr2):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    lib().spaa_adam_step_dev(_p(param), _p(grad), _p(m), _p(v), param.numel(), _p(seg_end), _p(seg_lr), _p(seg_wd), seg_end.numel(),
                             beta1, beta2, eps, _p(bias_corr2), float(grad_scale), _stream()); _count()


# ------------------------------------------------------------------------------------------------------------
# attack-loop updates
# ------------------------------------------------------------------------------------------------------------

def row_sqnorm(g: Tensor, sq: Tensor, x_for_clamp: Optional[Tensor] = None, lo: float = 0.0, hi: float = 1.0) -> Tensor:
    B = g.shape[0]
    n = g.numel() // B
    L = lib()
    ws = workspace("rownorm", L.spaa_rownorm_ws_bytes(B, n), g.device)
    L.spaa_row_sqnorm(_p(g), _p(x_for_clamp), lo, hi, B, n, _p(sq), _p(ws), _stream()); _count()
    return sq


def row_normalized_step(x: Tensor, g: Tensor, sq: Tensor, step2: Tensor, sel: Optional[Tensor] = None, *, use_clamp_mask: bool = False,
                        lo: float = 0.0, hi: float = 1.0, base: Optional[Tensor] = None, sum_out: Optional[Tensor] = None,
                        copy_dst: Optional[Tensor] = None, copy_sel: Optional[Tensor] = None) -> None:
    B = x.shape[0]
    n = x.numel() // B
    lib().spaa_row_normalized_step(_p(x), _p(g), _p(sq), _p(sel), _p(step2), int(use_clamp_mask), lo, hi, _p(base),
                                   _bstride(base, B) if base is not None else 0, _p(sum_out), _p(copy_dst), _p(copy_sel), B, n,
                                   _stream()); _count()


def masked_copy_rows(dst: Tensor, src: Tensor, sel: Tensor) -> None:
    B = dst.shape[0]
    lib().spaa_masked_copy_rows(_p(dst), _p(src), _p(sel), B, dst.numel() // B, _stream()); _count()


def select_cotangent(g0: Tensor, g1: Optional[Tensor], sel: Optional[Tensor], act: Optional[Tensor], mask_mode: int, out: Tensor) -> Tensor:
    B = g0.shape[0]
    lib().spaa_select_cotangent(_p(g0), _p(g1), _p(sel), _p(act), mask_mode, _p(out), B, g0.numel() // B, _stream()); _count()
    return out


def percal_project(base: Tensor, delta: Tensor, xq: Tensor, xsum: Optional[Tensor], l2sum: Tensor) -> None:
    B, _, H, W = delta.shape
    L = lib()
    ws = workspace("percal_project", L.spaa_percal_project_ws_bytes(B, H * W), delta.device)
    L.spaa_percal_project(_p(base), _bstride(base, B), _p(delta), _p(xq), _p(xsum), _p(l2sum), B, H * W, _p(ws), _stream()); _count()


def chan_l2(x: Tensor, ref: Tensor, sums: Tensor, *, c: float = 0.0, sel: Optional[Tensor] = None, apply_clamp_mask: bool = False,
            grad: Optional[Tensor] = None) -> None:
    B, _, H, W = x.shape
    L = lib()
    ws = workspace("chan_l2", L.spaa_chan_l2_ws_bytes(B, H * W), x.device)
    L.spaa_chan_l2_fwd_bwd(_p(x), _p(ref), _bstride(ref, B), B, H * W, float(c), _p(sel), int(apply_clamp_mask), _p(sums), _p(grad), _p(ws),
                           _stream()); _count()


def attack_masks(logits: Tensor, target: Tensor, targeted: bool, stats: Tensor, prjl2sum: Optional[Tensor], hw_cam: int, hw_prj: int,
                 w_prjl2: float, w_caml2: float, w_camde: float, d_thr: float, p_thresh: float, use_col: Tensor, succ: Tensor,
                 better: Tensor, col_loss: Tensor, best_col: Tensor) -> None:
    logits = _f32c(logits)
    B, ncls = logits.shape
    lib().spaa_attack_masks(_p(logits), ncls, _p(target), int(targeted), _p(stats), _p(prjl2sum), hw_cam, hw_prj, float(w_prjl2),
                            float(w_caml2), float(w_camde), float(d_thr), float(p_thresh), B, _p(use_col), _p(succ), _p(better),
                            _p(col_loss), _p(best_col), _stream()); _count()


def percal_masks(logits: Tensor, labels: Tensor, mode: int, margin: float, l2sum: Tensor, hw: int, d_thr: float, p_thresh: float,
                 stats: Tensor, isadv: Tensor, use_col: Tensor, better: Tensor, dis: Tensor, best_dis: Tensor) -> None:
    logits = _f32c(logits)
    B, ncls = logits.shape
    lib().spaa_percal_masks(_p(logits), ncls, _p(labels), mode, float(margin), _p(l2sum), hw, float(d_thr), float(p_thresh), _p(stats), B,
                            _p(isadv), _p(use_col), _p(better), _p(dis), _p(best_dis), _stream()); _count()


# ------------------------------------------------------------------------------------------------------------
# classifier pre-processing (fused crop + area resize + normalise)
# ------------------------------------------------------------------------------------------------------------

def _f3(v):
    return (ctypes.c_float * 3)(*[float(t) for t in v])


S2D_PAD_LO, S2D_PAD_HI, S2D_C = 2, 1, 16          # layout 2 of spaa_clf_preprocess_fwd (include/spaa_b200.h)


def clf_preprocess(img: Tensor, crop, out_hw, mean, std, channels_last: bool, s2d: bool = False) -> Tensor:
    """Returns the network input as a logical [B,3,h,w] tensor (NCHW, or channels-last memory when `channels_last`); with `s2d` the
    2x2 space-to-depth fold a classifier.S2DStem reads: logical [B,16,h/2+3,w/2+3] in channels_last memory."""
    img = _f32c(img)
    B, C, H, W = img.shape
    assert C == 3
    top, left, ch, cw = crop
    if s2d:
        assert out_hw[0] % 2 == 0 and out_hw[1] % 2 == 0
        pad = S2D_PAD_LO + S2D_PAD_HI
        out = torch.empty((B, out_hw[0] // 2 + pad, out_hw[1] // 2 + pad, S2D_C), dtype=torch.float32, device=img.device)
        lib().spaa_clf_preprocess_fwd(_p(img), B, H, W, top, left, ch, cw, out_hw[0], out_hw[1], _f3(mean), _f3(std), 2, _p(out), _stream()); _count()
        return out.permute(0, 3, 1, 2)
    if channels_last:
        out = torch.empty((B, out_hw[0], out_hw[1], 3), dtype=torch.float32, device=img.device)
    else:
        out = torch.empty((B, 3, out_hw[0], out_hw[1]), dtype=torch.float32, device=img.device)
    lib().spaa_clf_preprocess_fwd(_p(img), B, H, W, top, left, ch, cw, out_hw[0], out_hw[1], _f3(mean), _f3(std), int(channels_last), _p(out), _stream()); _count()
    return out.permute(0, 3, 1, 2) if channels_last else out


def clf_preprocess_bwd(dout: Tensor, img_hw, crop, mean, std, s2d: bool = False) -> Tensor:
    """dout: logical [B,3,h,w], NCHW-contiguous or channels-last (s2d: the gradient of the folded input, logical [B,16,h/2+3,w/2+3]); returns
    d/d(img) [B,3,H,W] (zero outside the crop)."""
    if dout.dtype != torch.float32:
        dout = dout.float()
    if s2d:
        assert dout.shape[1] == S2D_C
        if not _nhwc_dense(dout):
            dout = dout.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        B = dout.shape[0]
        pad = S2D_PAD_LO + S2D_PAD_HI
        oh, ow = 2 * (dout.shape[2] - pad), 2 * (dout.shape[3] - pad)
        top, left, ch, cw = crop
        dimg = torch.empty((B, 3, img_hw[0], img_hw[1]), dtype=torch.float32, device=dout.device)
        lib().spaa_clf_preprocess_bwd(_p(dout), B, img_hw[0], img_hw[1], top, left, ch, cw, oh, ow, _f3(std), 2, _p(dimg), _stream()); _count()
        return dimg
    nhwc = dout.is_contiguous(memory_format=torch.channels_last) and not dout.is_contiguous()
    if not nhwc and not dout.is_contiguous():
        dout = dout.contiguous()
    B, _, oh, ow = dout.shape
    top, left, ch, cw = crop
    dimg = torch.empty((B, 3, img_hw[0], img_hw[1]), dtype=torch.float32, device=dout.device)
    lib().spaa_clf_preprocess_bwd(_p(dout), B, img_hw[0], img_hw[1], top, left, ch, cw, oh, ow, _f3(std), int(nhwc), _p(dimg), _stream()); _count()
    return dimg


# ------------------------------------------------------------------------------------------------------------
# fused ReLU + max-pooling of the external classifier's first stage (channels_last fp32)
# ------------------------------------------------------------------------------------------------------------

def _nhwc_dense(x: Tensor) -> bool:
    """Logical [N,C,H,W] tensor stored densely as [N,H,W,C] (strides of size-1 dimensions are arbitrary in torch and ignored)."""
    N, C, H, W = x.shape
    want = (H * W * C, 1, W * C, C)
    return all(sz == 1 or st == w for sz, st, w in zip(x.shape, x.stride(), want))


def relu_maxpool_supported(x: Tensor, k: int, stride: int, pad: int) -> bool:
    """True when `x` is what spaa_relu_maxpool_nhwc_fwd reads in place: a CUDA fp32 [N,C,H,W] tensor in channels_last memory, C % 4 == 0."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        return False
    N, C, H, W = x.shape
    if C % 4 or not (1 <= k <= 15) or stride < 1 or pad < 0 or 2 * pad > k or H + 2 * pad < k or W + 2 * pad < k or max(N, H) > 65535:
        return False
    return _nhwc_dense(x) and x.data_ptr() % 16 == 0


def _bias_ok(bias: Optional[Tensor], x: Tensor) -> bool:
    return bias is None or (bias.is_cuda and bias.dtype == torch.float32 and bias.dim() == 1 and bias.shape[0] == x.shape[1] and bias.is_contiguous()
                            and bias.data_ptr() % 16 == 0)


def bias_act_supported(x: Tensor, bias: Optional[Tensor], res: Optional[Tensor]) -> bool:
    """True when spaa_bias_act_nhwc reads these operands in place: CUDA fp32 [N,C,H,W] in channels_last memory, C % 4 == 0."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0 and x.numel() > 0):
        return False
    if not (_nhwc_dense(x) and x.data_ptr() % 16 == 0 and _bias_ok(bias, x) and x.numel() // 4 < 2 ** 32):
        return False
    return res is None or (res.is_cuda and res.dtype == torch.float32 and res.shape == x.shape and _nhwc_dense(res) and res.data_ptr() % 16 == 0)


def bias_act_nhwc(x: Tensor, bias: Optional[Tensor], res: Optional[Tensor], relu: bool) -> Tensor:
    """relu?(x + bias[c] + res) in one pass; a new tensor with x's (channels_last) layout."""
    if not bias_act_supported(x, bias, res):
        raise RuntimeError("bias_act_nhwc needs CUDA fp32 channels_last tensors with C % 4 == 0")
    y = torch.empty_like(x)          # preserve_format: dense channels_last like x
    if not _nhwc_dense(y):
        y = torch.empty(x.shape, dtype=x.dtype, device=x.device).contiguous(memory_format=torch.channels_last)
    lib().spaa_bias_act_nhwc(_p(x), _p(bias), _p(res), x.numel(), x.shape[1], int(bool(relu)), _p(y), _stream()); _count()
    return y


def relu_maxpool_nhwc(x: Tensor, k: int, stride: int, pad: int, relu: bool, bias: Optional[Tensor] = None):
    """max_pool2d(relu(x + bias) if relu else x + bias, k, stride, pad) for a channels_last fp32 tensor.  Returns (y, idx): y logical
    [N,C,Ho,Wo] in channels_last memory, idx uint8 [N,Ho,Wo,C] (the tap that was selected; 255: none) for relu_maxpool_nhwc_bwd."""
    if not relu_maxpool_supported(x, k, stride, pad) or not _bias_ok(bias, x):
        raise RuntimeError("relu_maxpool_nhwc needs a CUDA fp32 channels_last tensor with C % 4 == 0")
    N, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device=x.device)
    idx = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    lib().spaa_relu_maxpool_nhwc_fwd(_p(x), _p(bias), N, H, W, C, k, stride, pad, Ho, Wo, int(bool(relu)), _p(y), _p(idx), _stream()); _count()
    return y.permute(0, 3, 1, 2), idx


def relu_maxpool_nhwc_bwd(dy: Tensor, idx: Tensor, in_hw, k: int, stride: int, pad: int) -> Tensor:
    """Adjoint of relu_maxpool_nhwc: dy logical [N,C,Ho,Wo] -> dx logical [N,C,H,W] (channels_last memory)."""
    _need_cuda(dy, idx)
    if dy.dtype != torch.float32:
        dy = dy.float()
    N, C, Ho, Wo = dy.shape
    if not _nhwc_dense(dy):
        dy = dy.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    H, W = in_hw
    assert idx.shape == (N, Ho, Wo, C) and idx.dtype == torch.uint8 and idx.is_contiguous()
    dx = torch.empty((N, H, W, C), dtype=torch.float32, device=dy.device)
    lib().spaa_relu_maxpool_nhwc_bwd(_p(dy), _p(idx), N, H, W, C, k, stride, pad, Ho, Wo, _p(dx), _stream()); _count()
    return dx.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------------
# device guard: every wrapper above launches on `torch.cuda.current_stream()` of the CURRENT device, while the reference API takes
# `device` / `device_ids` arguments (spaa(..., device='cuda:1'), cfg.device) independent of it.  Each public op therefore runs with the
# device of its tensor operands made current, and refuses operands that live on different devices.
# ------------------------------------------------------------------------------------------------------------

def _operand_device(args, kwargs):
    dev = None
    for a in list(args) + list(kwargs.values()):
        if torch.is_tensor(a) and a.is_cuda:
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise ValueError(f"spaa_b200 op called with tensors on different devices ({dev} and {a.device})")
    return dev


NVTX = os.environ.get("SPAA_NVTX", "0") not in ("", "0")      # one NVTX range per op wrapper (= per kernel family) for nsys / ncu --nvtx timelines


def _device_guarded(fn):
    import functools
    name = "spaa." + fn.__name__

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _operand_device(args, kwargs)
        if NVTX:
            torch.cuda.nvtx.range_push(name)
        try:
            if dev is None or dev.index == torch.cuda.current_device():
                return fn(*args, **kwargs)
            with torch.cuda.device(dev):
                return fn(*args, **kwargs)
        finally:
            if NVTX:
                torch.cuda.nvtx.range_pop()
    return wrapper


for _name in ("rgb2lab", "rgb2lab_bwd", "de2000", "de2000_bwd", "color_loss", "tps_grid", "coarse_grid", "coarse_grid_bwd", "grid_finish",
              "grid_finish_bwd", "grid_sample", "grid_sample_packed", "pack_nhwc16", "select_cotangent_packed", "grid_sample_bwd_input",
              "grid_sample_bwd_gather", "grid_sample_bwd_grid", "conv_forward", "conv_backward_data", "conv_backward_weight", "channel_sum",
              "ssim_l1", "adam_step", "adam_step_dev", "row_sqnorm", "row_normalized_step", "masked_copy_rows", "select_cotangent",
              "percal_project", "chan_l2", "attack_masks", "percal_masks", "clf_preprocess", "clf_preprocess_bwd", "bias_act_nhwc",
              "relu_maxpool_nhwc", "relu_maxpool_nhwc_bwd"):
    globals()[_name] = _device_guarded(globals()[_name])
WarpAdjoint.__init__ = _device_guarded(WarpAdjoint.__init__)
del _name
