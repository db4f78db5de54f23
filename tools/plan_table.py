"""Launch plans of the halo convolution kernel for every ShadingNet layer at the BASELINE shapes (B=32, 240x320), forward and backward-data:
host arithmetic only (spaa_conv_tc_plan), runs without a GPU.  usage: python tools/plan_table.py [B]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from spaa_b200 import ops  # noqa: E402
from spaa_b200._lib import ConvDesc, lib  # noqa: E402
from spaa_b200.models import _stack_specs  # noqa: E402

B = int(sys.argv[1]) if __name__ == "__main__" and len(sys.argv) > 1 else 32
H, W = 240, 320


def act(c, h, w, dt=torch.float16):
    return torch.empty((B, c, h, w), dtype=dt, device="meta").contiguous(memory_format=torch.channels_last) if dt != torch.float32 else \
        torch.empty((B, c, h, w), dtype=dt, device="meta")


def plan(d, add, mask, mask2):
    p = (ctypes.c_int32 * 12)()
    rc = lib().cdll.spaa_conv_tc_plan(ctypes.byref(d), int(add), int(mask), int(mask2), p)
    return list(p) if rc == 0 else None


def fwd(spec, cin_t, hin, win, add=False, planar=False):
    ho, wo = spec.out_hw(hin, win)
    d = ConvDesc()
    d.Cin, d.Cout, d.KH, d.KW = cin_t, spec.cout, spec.k, spec.k
    if spec.kind == "conv":
        d.stride, d.up, d.pad_h, d.pad_w, d.flip = spec.stride, 1, spec.pad, spec.pad, 0
    else:
        d.stride, d.up, d.flip = 1, spec.stride, 1
        d.pad_h = d.pad_w = spec.k - 1 - spec.pad
    out = act(spec.cout, ho, wo, torch.float32 if planar else torch.float16)
    ops._fill_desc(d, act(cin_t, hin, win), out, out if add else None, None)
    return plan(d, add, False, False), (ho, wo)


def bwd(spec, cout_t, hin, win, add=False, mask=False, mask2=False, planar=False, cin_out=None):
    ho, wo = spec.out_hw(hin, win)
    cin = cin_out or spec.cin
    d = ConvDesc()
    d.Cin, d.Cout, d.KH, d.KW = cout_t, cin, spec.k, spec.k
    if spec.kind == "conv":
        d.stride, d.up, d.flip = 1, spec.stride, 1
        d.pad_h = d.pad_w = spec.k - 1 - spec.pad
    else:
        d.stride, d.up, d.pad_h, d.pad_w, d.flip = spec.stride, 1, spec.pad, spec.pad, 0
    out = act(cin, hin, win, torch.float32 if planar else torch.bfloat16)
    ops._fill_desc(d, act(cout_t, ho, wo, torch.bfloat16), out, out if add else None, out if (mask or mask2) else None)
    d.mask_mode = ops.MASK_POS if mask else ops.MASK_NONE
    return plan(d, add, mask, mask2)


def layer_plans(batch=None, hw=None, variant="shading"):
    """[(layer, plan or None)] for the 27 tensor-core launches of one ShadingNetSPAA (variant 'shading') / CompenNet ('compen') forward +
    backward-data pass at `batch` images of `hw` pixels (default: the BASELINE attack shapes, 32 x 240 x 320)."""
    global B, H, W
    if batch is not None:
        B = batch
    if hw is not None:
        H, W = hw
    sp = _stack_specs(variant, 6)
    rows = []
    names = "ctas eg nbuf sa sb S dbuf pair smem tiles threads resident".split()
    r, _ = fwd(sp["conv1_s"], 16, H, W); rows.append(("conv1_s f", r))
    r, _ = fwd(sp["conv2_s"], 32, H // 2, W // 2); rows.append(("conv2_s f", r))
    r, _ = fwd(sp["conv3_s"], 64, H // 4, W // 4); rows.append(("conv3_s f", r))
    r, _ = fwd(sp["conv4_s"], 128, H // 4, W // 4); rows.append(("conv4_s f", r))
    r, _ = fwd(sp["conv1"], 16, H, W, add=True); rows.append(("conv1 f", r))
    r, _ = fwd(sp["skipConv2"], 32, H // 2, W // 2); rows.append(("skipConv2 f", r))
    r, _ = fwd(sp["conv2"], 32, H // 2, W // 2, add=True); rows.append(("conv2 f", r))
    r, _ = fwd(sp["skipConv3"], 64, H // 4, W // 4); rows.append(("skipConv3 f", r))
    r, _ = fwd(sp["conv3"], 64, H // 4, W // 4, add=True); rows.append(("conv3 f", r))
    r, _ = fwd(sp["conv4"], 128, H // 4, W // 4, add=True); rows.append(("conv4 f", r))
    r, _ = fwd(sp["conv5"], 256, H // 4, W // 4, add=True); rows.append(("conv5 f", r))
    r, _ = fwd(sp["transConv1"], 128, H // 4, W // 4, add=True); rows.append(("transConv1 f", r))
    r, _ = fwd(sp["transConv2"], 64, H // 2, W // 2); rows.append(("transConv2 f", r))
    r, _ = fwd(sp["conv6"], 32, H, W, add=True, planar=True); rows.append(("conv6 f", r))
    rows.append(("conv6 b", bwd(sp["conv6"], 16, H, W, mask=True)))
    rows.append(("transConv2 b", bwd(sp["transConv2"], 32, H // 2, W // 2, mask=True)))
    rows.append(("transConv1 b", bwd(sp["transConv1"], 64, H // 4, W // 4, mask=True)))
    rows.append(("conv5 b", bwd(sp["conv5"], 128, H // 4, W // 4, mask=True, mask2=True)))
    rows.append(("conv4 b", bwd(sp["conv4"], 256, H // 4, W // 4, mask=True)))
    rows.append(("conv3 b", bwd(sp["conv3"], 128, H // 4, W // 4)))
    rows.append(("skipConv3 b", bwd(sp["skipConv3"], 128, H // 4, W // 4, add=True, mask=True)))
    rows.append(("conv2 b", bwd(sp["conv2"], 64, H // 2, W // 2)))
    rows.append(("skipConv2 b", bwd(sp["skipConv2"], 64, H // 2, W // 2, add=True, mask=True)))
    rows.append(("conv1 b", bwd(sp["conv1"], 32, H, W, planar=True, cin_out=3)))
    rows.append(("conv3_s b", bwd(sp["conv3_s"], 128, H // 4, W // 4, add=True, mask=True)))
    rows.append(("conv2_s b", bwd(sp["conv2_s"], 64, H // 2, W // 2, add=True, mask=True)))
    rows.append(("conv1_s b", bwd(sp["conv1_s"], 32, H, W, planar=True, cin_out=6)))
    return rows, names


if __name__ == "__main__":
    rows, names = layer_plans()
    print("%-14s " % "layer" + " ".join("%7s" % n for n in names))
    for n, r in rows:
        print("%-14s " % n + (" ".join("%7d" % v for v in r) if r else "unsupported by the halo kernel"))
