"""PCNet / WarpingNet / ShadingNetSPAA / CompenNet / CompenNet++ with the reference's module API
(/root/reference/src/python/models.py:11-346) and state-dict layout, executed by the sm_100a kernels of
libspaa_b200.so.

Each network is ONE autograd node: its forward runs a hand-scheduled chain of fused conv kernels (bias, residual add,
ReLU / clamp in the epilogue) and keeps the activations; its backward runs the matching chain of backward-data
kernels with the ReLU masks and skip-connection sums fused into their epilogues, plus the backward-weight kernels
when parameters need gradients.  The nn.Conv2d / nn.ConvTranspose2d sub-modules only hold parameters (so reference
checkpoints load unchanged); their own forward is never called.

CUDA only -- calling a module with CPU tensors raises.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn as nn

from . import ops
from . import pytorch_tps
from .ops import ConvSpec, EPI_RELU, EPI_LEAKY01, EPI_CLAMP_MAX1, EPI_ADD_AFTER_ACT, MASK_POS, MASK_LEAKY01, MASK_OPEN01

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# shared conv-stack engine for ShadingNetSPAA and CompenNet (same topology, models.py:74-94 and :280-303)
# --------------------------------------------------------------------------------------------------------------

def _stack_specs(variant: str, surf_ch: int) -> Dict[str, ConvSpec]:
    shading = variant == "shading"
    return {
        "conv1": ConvSpec("conv", 3, 32, 3, 2, 1), "conv2": ConvSpec("conv", 32, 64, 3, 2, 1),
        "conv3": ConvSpec("conv", 64, 128, 3, 1, 1), "conv4": ConvSpec("conv", 128, 256, 3, 1, 1),
        "conv5": ConvSpec("conv", 256, 128, 3, 1, 1),
        "conv1_s": ConvSpec("conv", surf_ch, 32, 3, 2, 1), "conv2_s": ConvSpec("conv", 32, 64, 3, 2, 1),
        "conv3_s": ConvSpec("conv", 64, 128, 3, 1, 1), "conv4_s": ConvSpec("conv", 128, 256, 3, 1, 1),
        "transConv1": ConvSpec("convT", 128, 64, 3, 2, 1, 1) if shading else ConvSpec("convT", 128, 64, 2, 2, 0),
        "transConv2": ConvSpec("convT", 64, 32, 2, 2, 0), "conv6": ConvSpec("conv", 32, 3, 3, 1, 1),
        "skipConv1.0": ConvSpec("conv", 3, 3, 1, 1, 0) if shading else ConvSpec("conv", 3, 3, 3, 1, 1),
        "skipConv1.2": ConvSpec("conv", 3, 3, 3, 1, 1), "skipConv1.4": ConvSpec("conv", 3, 3, 3, 1, 1),
        "skipConv2": ConvSpec("conv", 32, 64, 1, 1, 0),
        "skipConv3": ConvSpec("conv", 64, 128, 3, 1, 1) if shading else ConvSpec("conv", 64, 128, 1, 1, 0),
    }


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(device) -> "torch.cuda.Stream":
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[idx]


def _get(net: nn.Module, name: str) -> nn.Module:
    m = net
    for part in name.split("."):
        m = m[int(part)] if part.isdigit() else getattr(m, part)
    return m


class _Stack:
    """Forward / backward schedules of the 17-conv stack.  `net` supplies parameters and specs."""

    @staticmethod
    def act_dtype(net):
        """Storage type of the forward activations: fp32 NCHW (exact CUDA-core path) or 16-bit NHWC (tcgen05 path)."""
        return {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16, "bf16x3": torch.bfloat16}[getattr(net, "precision", "fp32")]

    @staticmethod
    def split(net) -> bool:
        """'bf16x3': every activation / gradient tensor of the stack carries three bf16 parts per logical channel (ops.conv_forward, split=True)."""
        return getattr(net, "precision", "fp32") == "bf16x3"

    @staticmethod
    def grad_dtype(net):
        """Storage type of the backward activations: bf16 on both tensor-core modes (gradients need the exponent range)."""
        return torch.float32 if getattr(net, "precision", "fp32") == "fp32" else torch.bfloat16

    @staticmethod
    def surface_branch(net, surf: Tensor, packed: Optional[Tensor] = None, bf16_copy: Optional[dict] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        sp = net._specs
        ad = _Stack.act_dtype(net)
        s3 = _Stack.split(net)
        W = lambda n: (_get(net, n).weight, _get(net, n).bias)
        if s3 and packed is None:
            packed = ops.pack_nhwc16(surf, None, ad, split=True)               # surface image(s) alone (simplify(), no-rough models): channels 0..Cs-1
            r1s = ops.conv_forward(sp["conv1_s"], packed, *W("conv1_s"), epi=EPI_RELU, cin_offset=0, split=True)
        elif packed is not None:          # [x | s | x*s | 0] 16-channel NHWC: the surface features start at channel 3
            r1s = ops.conv_forward(sp["conv1_s"], packed, *W("conv1_s"), epi=EPI_RELU, cin_offset=3, split=s3, bf16_copy=bf16_copy)
        else:
            r1s = ops.conv_forward(sp["conv1_s"], surf, *W("conv1_s"), epi=EPI_RELU, out_dtype=ad)
        r2s = ops.conv_forward(sp["conv2_s"], r1s, *W("conv2_s"), epi=EPI_RELU, split=s3, bf16_copy=bf16_copy)
        r3s = ops.conv_forward(sp["conv3_s"], r2s, *W("conv3_s"), epi=EPI_RELU, split=s3, bf16_copy=bf16_copy)
        r4s = ops.conv_forward(sp["conv4_s"], r3s, *W("conv4_s"), epi=EPI_RELU, split=s3)
        return r1s, r2s, r3s, r4s

    @staticmethod
    def skip1(net, skip_in: Tensor):
        sp = net._specs
        W = lambda n: (_get(net, n).weight, _get(net, n).bias)
        t1 = ops.conv_forward(sp["skipConv1.0"], skip_in, *W("skipConv1.0"), epi=EPI_RELU)
        t2 = ops.conv_forward(sp["skipConv1.2"], t1, *W("skipConv1.2"), epi=EPI_RELU)
        res1 = ops.conv_forward(sp["skipConv1.4"], t2, *W("skipConv1.4"), epi=EPI_RELU)
        return t1, t2, res1

    @staticmethod
    def forward(net, x: Optional[Tensor], surf: Optional[Tensor], skip_in: Optional[Tensor], *, surf_acts=None, skip_acts=None,
                packed: Optional[Tensor] = None, bf16_copies: bool = False) -> Tuple[Tensor, dict]:
        """x [B,3,H,W]; surf [B or 1,Cs,H,W] (ignored when surf_acts given); skip_in [B or 1,3,H,W] (ignored when
        skip_acts given).  `packed`: the 16-channel NHWC tensor [x | s | x*s | 0] from ops.grid_sample_packed replaces x and
        surf on the tensor-core path.  Returns (out, saved activations)."""
        sp = net._specs
        W = lambda n: (_get(net, n).weight, _get(net, n).bias)
        if packed is None and x is not None and x.dtype == torch.float32 and _Stack.act_dtype(net) != torch.float32 and x.shape[1] == 3 and \
                (surf_acts is not None or (surf is not None and surf.dtype == torch.float32 and surf.shape[1] <= 6)):
            # images from the nn.Module API: one pass writes [x | surf | 0] as the 16-channel NHWC operand of the tensor-core conv1 / conv1_s
            packed = ops.pack_nhwc16(x, None if surf_acts is not None else surf, _Stack.act_dtype(net), split=_Stack.split(net))
        s3 = _Stack.split(net)
        S: dict = {"x": x, "surf": surf, "skip_in": skip_in, "packed": packed}
        # 'fp16' training: every activation that is the input of a layer is also written as bf16 by the epilogue that produces it (the
        # backward-weight kernel's operand format); backward() finds the copies in S["bf16_of"] by data pointer
        bc = {} if (bf16_copies and _Stack.act_dtype(net) == torch.float16 and not s3) else None
        if bc is not None:
            S["bf16_of"] = bc
        if surf_acts is None:
            surf_acts = _Stack.surface_branch(net, surf, packed, bf16_copy=bc)
            S["surf_own"] = True
        r1s, r2s, r3s, r4s = surf_acts
        if skip_acts is None:
            skip_acts = _Stack.skip1(net, skip_in)
            S["skip_own"] = True
        t1, t2, res1 = skip_acts
        if packed is not None:
            x1 = ops.conv_forward(sp["conv1"], packed, *W("conv1"), add=r1s, epi=EPI_RELU, cin_offset=0, split=s3, bf16_copy=bc)
        else:
            x1 = ops.conv_forward(sp["conv1"], x, *W("conv1"), add=r1s, epi=EPI_RELU, out_dtype=r1s.dtype)
        res2 = ops.conv_forward(sp["skipConv2"], x1, *W("skipConv2"), split=s3)
        x2 = ops.conv_forward(sp["conv2"], x1, *W("conv2"), add=r2s, epi=EPI_RELU, split=s3, bf16_copy=bc)
        res3 = ops.conv_forward(sp["skipConv3"], x2, *W("skipConv3"), split=s3)
        x3 = ops.conv_forward(sp["conv3"], x2, *W("conv3"), add=r3s, epi=EPI_RELU, split=s3, bf16_copy=bc)
        x4 = ops.conv_forward(sp["conv4"], x3, *W("conv4"), add=r4s, epi=EPI_RELU, split=s3, bf16_copy=bc)
        x5 = ops.conv_forward(sp["conv5"], x4, *W("conv5"), add=res3, epi=EPI_RELU, split=s3, bf16_copy=bc)
        x6 = ops.conv_forward(sp["transConv1"], x5, *W("transConv1"), add=res2, epi=EPI_RELU, split=s3, bf16_copy=bc)
        x7 = ops.conv_forward(sp["transConv2"], x6, *W("transConv2"), epi=EPI_RELU, split=s3, bf16_copy=bc)
        out = ops.conv_forward(sp["conv6"], x7, *W("conv6"), add=res1, epi=EPI_RELU | EPI_CLAMP_MAX1, out_dtype=torch.float32, split=s3)
        S.update(r1s=r1s, r2s=r2s, r3s=r3s, r4s=r4s, t1=t1, t2=t2, res1=res1, x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, x6=x6, x7=x7, out=out)
        return out, S

    @staticmethod
    def backward(net, S: dict, d_pre6: Optional[Tensor], *, need_dx: bool = True, surf_grad_channels: Optional[Tuple[int, int]] = None,
                 need_dskip: bool = False, param_grads: Optional[Dict[str, Tensor]] = None, d_pre6_packed: Optional[Tensor] = None):
        """d_pre6 = d(loss)/d(conv6 pre-activation) = dout * [0 < out < 1]: fp32 [B,3,H,W], or on the tensor-core path the
        and/or (tensor-core path) `d_pre6_packed`, the same values as a zero-padded 16-channel bf16 NHWC tensor from
        ops.select_cotangent_packed, which feeds the backward-data chain (the fp32 form is still needed for parameter gradients).
        Returns (dx, dsurf[:, lo:hi] or None, dskip_in or None); accumulates parameter gradients into `param_grads`
        ({"conv1.weight": fp32 tensor, ...}, zero-filled by the caller) when given."""
        sp = net._specs
        Wt = lambda n: _get(net, n).weight
        pg = param_grads
        s3 = _Stack.split(net)
        hw = lambda t: (t.shape[2], t.shape[3])
        B = (d_pre6 if d_pre6 is not None else d_pre6_packed).shape[0]

        # 'fp16' mode: bf16 copies of the fp16 forward activations, one per tensor (x1, x2 feed two layers each): written by the forward epilogues
        # (forward(bf16_copies=True)), else made here on first use
        bf16_of = S.get("bf16_of") if S.get("bf16_of") is not None else {}

        bias_jobs: list = []                           # (cotangent, bias gradient) of every tensor-core layer: summed in ONE launch at the end
        wg_scratch = ops.wgrad_scratch(S["out"].device) if (pg is not None and ops.WGRAD_SCRATCH and S["out"].is_cuda) else None

        # Parameter gradients run on a second stream ($SPAA_WGRAD_STREAM=0: same stream).  They are off the critical path (nothing reads them before the
        # optimiser), while the backward-data chain is strictly sequential: forked, the tail of every persistent kernel (CTAs that ran out of tiles, the
        # accumulator flush of the backward-weight kernel) overlaps with the other chain's kernels -- measured 8 750 -> 8 965 img/s at batch 24, same
        # job.  Joined before this function returns (the data-parallel gradient exchange starts there); works inside the captured step (fork / join).
        # (not in the split-precision mode: both chains are MMA-bound there and only slow each other down -- 1 483 img/s on one stream, 1 400 on two)
        side = _side_stream(S["out"].device) if (pg is not None and S["out"].is_cuda and not s3 and os.environ.get("SPAA_WGRAD_STREAM", "1") != "0") else None

        def wgrad(name, inp, dy, x_offset=0):
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(dy.device))
                with torch.cuda.stream(side):
                    _wgrad(name, inp, dy, x_offset)
            else:
                _wgrad(name, inp, dy, x_offset)

        def _wgrad(name, inp, dy, x_offset=0):
            if pg is not None and (name + ".weight") in pg:
                if inp.shape[0] == 1 and B > 1:
                    inp = inp.expand(B, -1, -1, -1)
                if inp.dtype == torch.float16 and dy.dtype == torch.bfloat16 and inp.is_contiguous(memory_format=torch.channels_last) and inp.numel() % 8 == 0:
                    # tcgen05 kind::f16 wants both backward-weight operands in one format: re-round the fp16 activation to bf16 (weight gradient only)
                    key = inp.data_ptr()
                    if key not in bf16_of:
                        bf16_of[key] = ops.half_to_bf16(inp)
                    inp = bf16_of[key]
                ops.conv_backward_weight(sp[name], inp, dy, pg[name + ".weight"], pg.get(name + ".bias"), x_offset=x_offset, split=s3 and inp.dtype != torch.float32,
                                         defer_bias=bias_jobs, scratch=wg_scratch)

        def packed_input():
            """[x | surf | 0] as one zero-padded 16-channel NHWC tensor in the gradient dtype: the X operand of the tensor-core
            backward-weight kernel for conv1 (channels 0-2) and conv1_s (channels 3..3+Cs); built once per backward."""
            if "packed_bw" not in S:
                if S.get("packed") is not None and S["packed"].dtype == gdt:
                    S["packed_bw"] = S["packed"]
                else:
                    S["packed_bw"] = ops.pack_nhwc16(S["x"], S["surf"], gdt)
            return S["packed_bw"]

        surf_live = S.get("surf_own", False) and (surf_grad_channels is not None or pg is not None)
        if surf_live and S["r1s"].shape[0] != B:
            raise RuntimeError("gradients through a batch-broadcast surface branch are not supported; expand `s` to the batch")
        x7, x6, x5, x4, x3, x2, x1 = S["x7"], S["x6"], S["x5"], S["x4"], S["x3"], S["x2"], S["x1"]
        gdt = _Stack.grad_dtype(net)
        in_hw = hw(S["packed"]) if S.get("packed") is not None else hw(S["x"])
        d7 = ops.conv_backward_data(sp["conv6"], d_pre6_packed if d_pre6_packed is not None else d_pre6, Wt("conv6"), hw(x7), mask=x7,
                                    mask_mode=MASK_POS, out_dtype=gdt, split=s3)
        if d_pre6_packed is not None and (d_pre6_packed.dtype == x7.dtype or (x7.dtype == torch.float16 and d_pre6_packed.dtype == torch.bfloat16)):
            wgrad("conv6", x7, d_pre6_packed)                       # tensor-core backward-weight: padded 16-channel cotangent (3 real)
        elif d_pre6 is not None:
            wgrad("conv6", x7, d_pre6)
        d6 = ops.conv_backward_data(sp["transConv2"], d7, Wt("transConv2"), hw(x6), mask=x6, mask_mode=MASK_POS, split=s3)
        wgrad("transConv2", x6, d7)
        d5 = ops.conv_backward_data(sp["transConv1"], d6, Wt("transConv1"), hw(x5), mask=x5, mask_mode=MASK_POS, split=s3)
        wgrad("transConv1", x5, d6)
        d4s = torch.empty_like(x4, dtype=gdt) if surf_live else None
        d4 = ops.conv_backward_data(sp["conv5"], d5, Wt("conv5"), hw(x4), mask=x4, mask_mode=MASK_POS,
                                    mask2=S["r4s"] if surf_live else None, out2=d4s, split=s3)
        wgrad("conv5", x4, d5)
        d3 = ops.conv_backward_data(sp["conv4"], d4, Wt("conv4"), hw(x3), mask=x3, mask_mode=MASK_POS, split=s3)
        wgrad("conv4", x3, d4)
        t2_ = ops.conv_backward_data(sp["conv3"], d3, Wt("conv3"), hw(x2), split=s3)
        wgrad("conv3", x2, d3)
        d2 = ops.conv_backward_data(sp["skipConv3"], d5, Wt("skipConv3"), hw(x2), add=t2_, mask=x2, mask_mode=MASK_POS, split=s3)
        wgrad("skipConv3", x2, d5)
        t1_ = ops.conv_backward_data(sp["conv2"], d2, Wt("conv2"), hw(x1), split=s3)
        wgrad("conv2", x1, d2)
        d1 = ops.conv_backward_data(sp["skipConv2"], d6, Wt("skipConv2"), hw(x1), add=t1_, mask=x1, mask_mode=MASK_POS, split=s3)
        wgrad("skipConv2", x1, d6)
        dx = None
        if need_dx:
            dx = ops.conv_backward_data(sp["conv1"], d1, Wt("conv1"), in_hw, out_dtype=torch.float32, split=s3)
        tc_bw = pg is not None and d1.dtype in (torch.bfloat16, torch.float16) and d1.dtype == gdt and (S["x"] is not None or S.get("packed") is not None)
        if tc_bw:
            wgrad("conv1", packed_input(), d1, 0)
        elif S["x"] is not None:
            wgrad("conv1", S["x"], d1)
        dsurf = None
        if surf_live:
            r3s, r2s, r1s = S["r3s"], S["r2s"], S["r1s"]
            d3s = ops.conv_backward_data(sp["conv4_s"], d4s, Wt("conv4_s"), hw(r3s), add=d3, mask=r3s, mask_mode=MASK_POS, split=s3)
            wgrad("conv4_s", r3s, d4s)
            d2s = ops.conv_backward_data(sp["conv3_s"], d3s, Wt("conv3_s"), hw(r2s), add=d2, mask=r2s, mask_mode=MASK_POS, split=s3)
            wgrad("conv3_s", r2s, d3s)
            d1s = ops.conv_backward_data(sp["conv2_s"], d2s, Wt("conv2_s"), hw(r1s), add=d1, mask=r1s, mask_mode=MASK_POS, split=s3)
            wgrad("conv2_s", r1s, d2s)
            if surf_grad_channels is not None:
                lo, hi = surf_grad_channels
                dsurf = ops.conv_backward_data(sp["conv1_s"], d1s, Wt("conv1_s")[:, lo:hi], in_hw, out_dtype=torch.float32, split=s3)
            if tc_bw:
                wgrad("conv1_s", packed_input(), d1s, 3)
            elif S["surf"] is not None:
                wgrad("conv1_s", S["surf"], d1s)
        dskip = None
        skip_params = pg is not None and "skipConv1.4.weight" in pg
        if S.get("skip_own", False) and (need_dskip or skip_params):
            t1, t2, res1 = S["t1"], S["t2"], S["res1"]
            d_res = d_pre6
            if res1.shape[0] != B:
                # skipConv1 ran ONCE on a surface image shared by the batch (training: the scene is expanded, train_network.py:297): its
                # activations and ReLU masks are the same for every sample and the branch is linear given the masks, so the parameter
                # gradients need only the batch SUM of the incoming gradient -- one reduction, then a B = 1 backward (24x less work)
                if need_dskip:
                    raise RuntimeError("the gradient wrt a batch-broadcast skipConv1 input is per sample; expand the input to get it")
                d_res = d_pre6.sum(0, keepdim=True)
            dr = ops.select_cotangent(d_res, None, None, res1, MASK_POS, torch.empty_like(d_res))
            dt2 = ops.conv_backward_data(sp["skipConv1.4"], dr, Wt("skipConv1.4"), hw(t2), mask=t2, mask_mode=MASK_POS)
            wgrad("skipConv1.4", t2, dr)
            dt1 = ops.conv_backward_data(sp["skipConv1.2"], dt2, Wt("skipConv1.2"), hw(t1), mask=t1, mask_mode=MASK_POS)
            wgrad("skipConv1.2", t1, dt2)
            if need_dskip:
                dskip = ops.conv_backward_data(sp["skipConv1.0"], dt1, Wt("skipConv1.0"), hw(S["skip_in"]))
            wgrad("skipConv1.0", S["skip_in"], dt1)
        if side is not None:
            with torch.cuda.stream(side):
                if bias_jobs:
                    ops.channel_sum_multi(bias_jobs)
                if wg_scratch is not None:
                    wg_scratch.flush()
            torch.cuda.current_stream(S["out"].device).wait_stream(side)
        else:
            if bias_jobs:
                ops.channel_sum_multi(bias_jobs)
            if wg_scratch is not None:
                wg_scratch.flush()
        return dx, dsurf, dskip


_STACK_PARAM_ORDER = ["conv1", "conv2", "conv3", "conv4", "conv5", "conv1_s", "conv2_s", "conv3_s", "conv4_s", "transConv1",
                      "transConv2", "conv6", "skipConv1.0", "skipConv1.2", "skipConv1.4", "skipConv2", "skipConv3"]


class _StackFn(torch.autograd.Function):
    """out = stack(x, surf, skip_in; params).  One autograd node for the whole 17-conv network."""

    @staticmethod
    def forward(ctx, net, x, surf, skip_in, use_cached_surf, *params):
        x = ops._f32c(x)
        surf_acts = skip_acts = None
        if use_cached_surf:
            surf_acts = tuple(t if t.dim() == 4 else t.unsqueeze(0) for t in (net.res1_s, net.res2_s, net.res3_s, net.res4_s))
        else:
            surf = ops._f32c(surf)
        if skip_in.dim() == 4 and skip_in.shape[0] > 1 and skip_in.stride(0) == 0 and not ctx.needs_input_grad[3]:
            skip_in = skip_in[:1]             # one image expanded over the batch: run skipConv1 once (see _Stack.backward)
        skip_in = ops._f32c(skip_in)
        with torch.no_grad():
            out, S = _Stack.forward(net, x, surf, skip_in, surf_acts=surf_acts, skip_acts=skip_acts, bf16_copies=any(ctx.needs_input_grad[5:]))
        ctx.net, ctx.S = net, S
        ctx.n_params = len(params)
        ctx.surf_given = not use_cached_surf
        return out

    @staticmethod
    def backward(ctx, dout):
        net, S = ctx.net, ctx.S
        need_x, need_surf, need_skip = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        names = [n for n, _ in net.named_parameters()]
        pneed = ctx.needs_input_grad[5:]
        pg = None
        direct = set()
        if any(pneed):
            # parameters managed by train_network.FlatAdam carry .grad views of one flat, already zeroed fp32 bucket: the backward-weight kernels
            # accumulate straight into them (they are `+=` kernels) and autograd is told there is nothing to add -- otherwise 46 zero-fills
            # and 46 accumulations per step exist only to move the same numbers once more
            pg = {}
            for (n, p), need in zip(net.named_parameters(), pneed):
                if not need:
                    continue
                if _flat_grad_view(p):
                    pg[n] = p.grad
                    direct.add(n)
                else:
                    pg[n] = torch.zeros_like(p, dtype=torch.float32)
        dout = ops._f32c(dout)
        d_pre6 = ops.select_cotangent(dout, None, None, S["out"], MASK_OPEN01, torch.empty_like(S["out"]))
        packed = None
        if _Stack.act_dtype(net) != torch.float32:
            Bq, _, Hq, Wq = dout.shape
            if _Stack.split(net):
                packed = ops.pack_nhwc16(d_pre6, None, torch.bfloat16, split=True)
            else:
                packed = torch.empty((Bq, 16, Hq, Wq), dtype=_Stack.grad_dtype(net), device=dout.device, memory_format=torch.channels_last)
                ops.select_cotangent_packed(dout, None, None, S["out"], MASK_OPEN01, packed)
        cs = S["surf"].shape[1] if (S["surf"] is not None and ctx.surf_given) else 0
        with torch.no_grad():
            dx, dsurf, dskip = _Stack.backward(net, S, d_pre6, need_dx=need_x, surf_grad_channels=(0, cs) if (need_surf and cs) else None,
                                               need_dskip=need_skip, param_grads=pg, d_pre6_packed=packed)
        ctx.S = None
        hook = getattr(net, "_grads_ready_hook", None)
        if hook is not None and pg is not None and len(direct) == len(pg):
            hook()                 # every parameter gradient of this stack now sits, complete, in the flat bucket (train_network.GradOverlap)
        pgr = tuple((pg.get(n) if (pg is not None and n not in direct) else None) for n in names)
        return (None, dx, dsurf, dskip, None) + pgr


def _flat_grad_view(p) -> bool:
    """True when `p.grad` is a live fp32 view that train_network.FlatAdam installed (and zero-fills before every backward)."""
    g = p.grad
    return bool(getattr(p, "_spaa_flat_grad", False)) and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape


def set_precision(model: nn.Module, precision: str) -> nn.Module:
    """'fp32': exact CUDA-core convolutions, fp32 NCHW activations (1e-5 parity mode).
    'bf16': tcgen05 tensor-core convolutions, bf16 NHWC activations and gradients, fp32 accumulation.
    'fp16': the same kernels with fp16 forward activations (3 more mantissa bits) and bf16 gradients.
    'bf16x3': fp32-accurate split-precision tensor-core mode -- every value as three bf16 parts (24 significand bits), every product as the six
              leading part products, fp32 accumulation in TMEM (include/spaa_b200.h, spaa_conv_desc.split); forward + backward-data."""
    if precision not in ("fp32", "bf16", "fp16", "bf16x3"):
        raise ValueError("precision must be 'fp32', 'bf16', 'fp16' or 'bf16x3'")
    for m in model.modules():
        if isinstance(m, (_ConvStackNet, WarpingNet)):
            m.precision = precision
    return model


class _ConvStackNet(nn.Module):
    """Common parameter container + engine access for ShadingNetSPAA and CompenNet."""
    precision = "fp32"

    def _build(self, variant: str, surf_ch: int):
        self._variant = variant
        self._specs = _stack_specs(variant, surf_ch)
        sp = self._specs
        mk = lambda s: (nn.Conv2d(s.cin, s.cout, s.k, s.stride, s.pad) if s.kind == "conv"
                        else nn.ConvTranspose2d(s.cin, s.cout, s.k, s.stride, s.pad, s.outpad))
        self.relu = nn.ReLU()
        for n in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv1_s", "conv2_s", "conv3_s", "conv4_s", "transConv1", "transConv2", "conv6"):
            setattr(self, n, mk(sp[n]))
        self.skipConv1 = nn.Sequential(mk(sp["skipConv1.0"]), self.relu, mk(sp["skipConv1.2"]), self.relu, mk(sp["skipConv1.4"]), self.relu)
        self.skipConv2 = mk(sp["skipConv2"])
        self.skipConv3 = mk(sp["skipConv3"])
        for n in ("res1_s", "res2_s", "res3_s", "res4_s"):     # surface-branch activations after simplify() (models.py:49-52)
            self.register_buffer(n, None)

        def _init(m):                                          # models.py:55-59: Conv2d only; ConvTranspose2d keeps the default init
            if type(m) == nn.Conv2d:
                nn.init.kaiming_normal_(m.weight)
        self.apply(_init)

    def simplify(self, s: Tensor):
        """models.py:62-71 / :268-277: cache the surface branch for a constant surface image."""
        with torch.no_grad():
            r = _Stack.surface_branch(self, ops._f32c(s))
        self.res1_s, self.res2_s, self.res3_s, self.res4_s = (t.squeeze() for t in r)

    def _run(self, x: Tensor, surf: Optional[Tensor], skip_in: Tensor) -> Tensor:
        cached = self.res1_s is not None
        params = [p for _, p in self.named_parameters()]
        return _StackFn.apply(self, x, None if cached else surf, skip_in, cached, *params)


class CompenNet(_ConvStackNet):
    """models.py:11-94."""

    def __init__(self):
        super().__init__()
        self.name = "CompenNet"
        self._build("compen", 3)

    def forward(self, x: Tensor, s: Tensor) -> Tensor:
        ops._need_cuda(x, s)
        return self._run(x, s, x)


class ShadingNetSPAA(_ConvStackNet):
    """models.py:214-303."""

    def __init__(self, use_rough: bool = True):
        super().__init__()
        self.use_rough = use_rough
        self.name = self.__class__.__name__ if use_rough else self.__class__.__name__ + "_no_rough"
        self._build("shading", 6 if use_rough else 3)

    def forward(self, x: Tensor, *argv: Tensor) -> Tensor:
        ops._need_cuda(x, *argv)
        surf = argv[0] if len(argv) == 1 else torch.cat(argv, 1)
        return self._run(x, surf, argv[0])


# --------------------------------------------------------------------------------------------------------------
# WarpingNet (models.py:98-185)
# --------------------------------------------------------------------------------------------------------------

_REFINE_SPECS = [ConvSpec("conv", 2, 32, 3, 2, 1), ConvSpec("conv", 32, 64, 3, 2, 1), ConvSpec("convT", 64, 32, 2, 2, 0),
                 ConvSpec("convT", 32, 2, 2, 2, 0)]
_REFINE_IDX = [0, 2, 4, 6]


class _CoarseGridFn(torch.autograd.Function):
    """grid_sample(affine_grid(affine; input size), tps_grid(theta; out size)) fused analytically -> planar [2,H,W]."""

    @staticmethod
    def forward(ctx, affine, theta, ctrl, in_hw, out_hw):
        ctx.save_for_backward(affine, theta, ctrl)
        ctx.in_hw, ctx.out_hw = in_hw, out_hw
        return ops.coarse_grid(affine, theta, ctrl, in_hw, out_hw)

    @staticmethod
    def backward(ctx, dgrid):
        affine, theta, ctrl = ctx.saved_tensors
        daff, dtheta = ops.coarse_grid_bwd(affine, theta, ctrl, ctx.in_hw, ctx.out_hw, dgrid)
        return daff.view_as(affine), dtheta.view_as(theta), None, None, None


class _RefineFn(torch.autograd.Function):
    """fine = clamp(refine_net(coarse) + coarse, -1, 1) on ONE copy of the grid (the reference runs the refinement net
    on B identical copies, models.py:172-176; the result and, by linearity, the gradients are the same)."""

    @staticmethod
    def forward(ctx, net, coarse, *params):
        g = coarse.unsqueeze(0)
        w = [net.grid_refine_net[i] for i in _REFINE_IDX]
        need_grad = any(p.requires_grad for p in params) or coarse.requires_grad
        # 16-bit TRAINING ('bf16' and 'fp16' modes): the three inner layers run on the tensor-core kernels (2-channel grid zero-padded to 16 NHWC channels, bf16 activations
        # and gradients, fp32 accumulation) -- on CUDA cores this B = 1 net cost 0.55 ms of a 4.1 ms step.  The last layer (LeakyReLU, residual
        # after the activation, fp32 output added to the fp32 coarse grid) and every frozen-model use (the attacks) stay exact fp32.
        tc = need_grad and getattr(net, "precision", "fp32") in ("bf16", "fp16") and ops.TC_ENABLED       # (bf16 inside, in both 16-bit training modes)
        gp = None
        if tc:
            gp = ops.pack_nhwc16(g, None, torch.bfloat16)
            a1 = ops.conv_forward(_REFINE_SPECS[0], gp, w[0].weight, w[0].bias, epi=EPI_RELU, out_dtype=torch.bfloat16)
        else:
            a1 = ops.conv_forward(_REFINE_SPECS[0], g, w[0].weight, w[0].bias, epi=EPI_RELU)
        a2 = ops.conv_forward(_REFINE_SPECS[1], a1, w[1].weight, w[1].bias, epi=EPI_RELU)
        a3 = ops.conv_forward(_REFINE_SPECS[2], a2, w[2].weight, w[2].bias, epi=EPI_RELU)
        # last layer: LeakyReLU(0.1) then + coarse (residual added after the activation)
        s = ops.conv_forward(_REFINE_SPECS[3], a3, w[3].weight, w[3].bias, add=g, epi=EPI_LEAKY01 | EPI_ADD_AFTER_ACT, out_dtype=torch.float32)
        pre4 = None
        if need_grad:
            pre4 = s - g            # leaky(pre): sign(pre) == sign(leaky(pre)) so it serves as the LeakyReLU mask
        fine = ops.grid_finish(s[0], None)
        ctx.net = net
        ctx.saved = (g, a1, a2, a3, s, pre4, gp)
        return fine

    @staticmethod
    def backward(ctx, dfine):
        net = ctx.net
        g, a1, a2, a3, s, pre4, gp = ctx.saved
        w = [net.grid_refine_net[i] for i in _REFINE_IDX]
        ds = ops.grid_finish_bwd(s[0], None, dfine).unsqueeze(0)                 # clamp backward
        d4 = ops.select_cotangent(ds, None, None, pre4, MASK_LEAKY01, torch.empty_like(ds))
        plist = [p for m in w for p in (m.weight, m.bias)]
        direct = [_flat_grad_view(p) for p in plist]
        grads = [p.grad if dflag else torch.zeros_like(p) for p, dflag in zip(plist, direct)]
        sc = ops.wgrad_scratch(g.device) if (ops.WGRAD_SCRATCH and g.is_cuda) else None      # one scatter launch for the four layers
        bias_jobs: list = []
        if gp is not None and a3.dtype == torch.bfloat16:
            # tensor-core training: the last layer's 2-channel cotangent goes to the tcgen05 kernels too, zero-padded to 16 bf16 NHWC channels like the
            # gradients of the three inner layers (on CUDA cores its backward-weight alone was 71 us of a 3.0 ms step)
            d4 = ops.pack_nhwc16(d4, None, torch.bfloat16)
        ops.conv_backward_weight(_REFINE_SPECS[3], a3, d4, grads[6], grads[7], scratch=sc, defer_bias=bias_jobs)
        d3 = ops.conv_backward_data(_REFINE_SPECS[3], d4, w[3].weight, a3.shape[2:], mask=a3, mask_mode=MASK_POS, out_dtype=a3.dtype)
        ops.conv_backward_weight(_REFINE_SPECS[2], a2, d3, grads[4], grads[5], scratch=sc, defer_bias=bias_jobs)
        d2 = ops.conv_backward_data(_REFINE_SPECS[2], d3, w[2].weight, a2.shape[2:], mask=a2, mask_mode=MASK_POS)
        ops.conv_backward_weight(_REFINE_SPECS[1], a1, d2, grads[2], grads[3], scratch=sc, defer_bias=bias_jobs)
        d1 = ops.conv_backward_data(_REFINE_SPECS[1], d2, w[1].weight, a1.shape[2:], mask=a1, mask_mode=MASK_POS)
        ops.conv_backward_weight(_REFINE_SPECS[0], gp if gp is not None else g, d1, grads[0], grads[1], scratch=sc, defer_bias=bias_jobs)
        if bias_jobs:
            ops.channel_sum_multi(bias_jobs)
        if sc is not None:
            sc.flush()
        dg = ops.conv_backward_data(_REFINE_SPECS[0], d1, w[0].weight, g.shape[2:], add=ds, out_dtype=torch.float32)      # + identity path
        ctx.saved = None
        return (None, dg[0]) + tuple(None if dflag else gr for gr, dflag in zip(grads, direct))


class _ClampGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coarse):
        ctx.save_for_backward(coarse)
        return ops.grid_finish(coarse, None)

    @staticmethod
    def backward(ctx, dfine):
        coarse, = ctx.saved_tensors
        return ops.grid_finish_bwd(coarse, None, dfine)


class _WarpFn(torch.autograd.Function):
    """out = grid_sample(x, grid) [* mask]; optionally sfeat = cat(s, out * s) written in the same pass."""

    @staticmethod
    def forward(ctx, x, grid, mask, s):
        x = ops._f32c(x)
        B, C = x.shape[:2]
        H, W = grid.shape[-2:]
        sfeat = None
        if s is not None:
            s = ops._f32c(s)
            sfeat = torch.empty((B, 2 * C, H, W), dtype=torch.float32, device=x.device)
            sfeat[:, :C] = s
            out = ops.grid_sample(x, grid, mask=mask, rough=s, out2=sfeat[:, C:])
        else:
            out = ops.grid_sample(x, grid, mask=mask)
        ctx.save_for_backward(x, grid, mask, s)
        if sfeat is None:
            return out
        return out, sfeat

    @staticmethod
    def backward(ctx, dout, dsfeat=None):
        x, grid, mask, s = ctx.saved_tensors
        C = x.shape[1]
        dout = ops._f32c(dout)
        d2 = None
        if dsfeat is not None and s is not None:
            dsfeat = ops._f32c(dsfeat)
            d2 = dsfeat[:, C:]
        dx = dgrid = ds = None
        if ctx.needs_input_grad[0]:
            dx = ops.grid_sample_bwd_input(dout, grid, x.shape[2:], mask=mask, dout2=d2, rough=s)
        if ctx.needs_input_grad[1]:
            dgrid = ops.grid_sample_bwd_grid(dout, x, grid, mask=mask, dout2=d2, rough=s)
            if grid.dim() == 4 and dgrid.dim() == 3:
                dgrid = dgrid.unsqueeze(0)
        if ctx.needs_input_grad[3] and dsfeat is not None:
            # sfeat = cat(s, out*s): d s = dsfeat[:, :C] + dsfeat[:, C:] * out   (rarely needed: s is data)
            out = ops.grid_sample(x, grid, mask=mask)
            ds = dsfeat[:, :C] + d2 * out
        return dx, dgrid, None, ds


class WarpingNet(nn.Module):
    precision = "fp32"          # set_precision: 'bf16' moves the refinement net's inner layers to the tensor-core kernels while training

    def __init__(self, grid_shape=(6, 6), out_size=(256, 256), with_refine=True):
        super().__init__()
        self.grid_shape = grid_shape
        self.out_size = tuple(out_size)
        self.with_refine = with_refine
        self.name = self.__class__.__name__ if with_refine else self.__class__.__name__ + "_without_refine"
        self.relu = nn.ReLU()
        self.leakyRelu = nn.LeakyReLU(0.1)
        self.register_buffer("fine_grid", None)
        self.affine_mat = nn.Parameter(torch.Tensor([1, 0, 0, 0, 1, 0]).view(-1, 2, 3))
        self.nctrl = self.grid_shape[0] * self.grid_shape[1]
        self.nparam = self.nctrl + 2
        self.register_buffer("ctrl_pts", pytorch_tps.uniform_grid(grid_shape).view(-1, 2))
        self.theta = nn.Parameter(torch.ones((1, self.nparam * 2), dtype=torch.float32).view(-1, self.nparam, 2) * 1e-3)

        def init_normal(m):                                    # models.py:124-126 (Conv2d only)
            if type(m) == nn.Conv2d:
                nn.init.normal_(m.weight, 0, 1e-4)

        if self.with_refine:
            self.grid_refine_net = nn.Sequential(
                nn.Conv2d(2, 32, 3, 2, 1), self.relu, nn.Conv2d(32, 64, 3, 2, 1), self.relu,
                nn.ConvTranspose2d(64, 32, 2, 2, 0), self.relu, nn.ConvTranspose2d(32, 2, 2, 2, 0), self.leakyRelu)
            self.grid_refine_net.apply(init_normal)
        else:
            self.grid_refine_net = None

    def set_affine(self, affine_vec):
        self.affine_mat.data = torch.Tensor(affine_vec).view(-1, 2, 3).to(self.affine_mat.device)

    def planar_grid(self, in_hw) -> Tensor:
        """The fine sampling grid as a planar [2,H,W] tensor shared by the whole batch (with autograd to the parameters)."""
        if self.fine_grid is not None:
            return self.fine_grid[0].permute(2, 0, 1).contiguous()
        coarse = _CoarseGridFn.apply(self.affine_mat, self.theta, self.ctrl_pts, tuple(in_hw), self.out_size)
        if self.with_refine:
            params = [p for m in (self.grid_refine_net[i] for i in _REFINE_IDX) for p in (m.weight, m.bias)]
            return _RefineFn.apply(self, coarse, *params)
        return _ClampGridFn.apply(coarse)

    def simplify(self, x: Tensor):
        """models.py:149-161: freeze the sampling grid (stored as 1xHxWx2 like the reference)."""
        with torch.no_grad():
            g = self.planar_grid(x.shape[2:])
        self.fine_grid = g.permute(1, 2, 0).unsqueeze(0).contiguous()

    def forward(self, x: Tensor) -> Tensor:
        ops._need_cuda(x)
        return _WarpFn.apply(x, self.planar_grid(x.shape[2:]), None, None)


class CompenNetPlusplus(nn.Module):
    """models.py:188-212."""

    def __init__(self, warping_net=None, compen_net=None):
        super().__init__()
        self.name = "CompenNet++"
        self.warping_net = copy.deepcopy(warping_net.module) if warping_net is not None else WarpingNet()
        self.compen_net = copy.deepcopy(compen_net.module) if compen_net is not None else CompenNet()

    def simplify(self, s):
        self.warping_net.simplify(s)
        self.compen_net.simplify(self.warping_net(s))

    def forward(self, x, s):
        ops._need_cuda(x, s)
        grid = self.warping_net.planar_grid(x.shape[2:])
        x = _WarpFn.apply(x, grid, None, None)
        s = _WarpFn.apply(s, grid, None, None)
        return self.compen_net(x, s)


class PCNet(nn.Module):
    """models.py:305-346."""

    def __init__(self, mask, warping_net=None, shading_net=None, fix_shading_net=False, use_mask=True, use_rough=True):
        super().__init__()
        self.name = self.__class__.__name__
        self.use_mask = use_mask
        self.use_rough = use_rough
        if not self.use_mask:
            self.name += "_no_mask"
        if not self.use_rough:
            self.name += "_no_rough"
        self.warping_net = copy.deepcopy(warping_net.module) if warping_net is not None else WarpingNet()
        self.shading_net = copy.deepcopy(shading_net.module) if shading_net is not None else ShadingNetSPAA()
        if self.use_mask:
            self.register_buffer("mask", mask.clone())
        for param in self.shading_net.parameters():
            param.requires_grad = not fix_shading_net

    def simplify(self, s):
        self.warping_net.simplify(s)
        self.shading_net.simplify(self.warping_net(s))

    def flat_mask(self) -> Optional[Tensor]:
        if not self.use_mask:
            return None
        return self.mask.to(torch.float32).reshape(-1).contiguous()

    def forward(self, x, s):
        ops._need_cuda(x, s)
        grid = self.warping_net.planar_grid(x.shape[2:])
        if s.shape[0] != x.shape[0]:
            s = s.expand(x.shape[0], -1, -1, -1)
        if self.use_rough:
            xw, sfeat = _WarpFn.apply(x, grid, self.flat_mask(), s)
            return self.shading_net._run(xw, sfeat, s)
        xw = _WarpFn.apply(x, grid, self.flat_mask(), None)
        return self.shading_net._run(xw, s, s)
