"""Per-layer relative error of the backward-data chain in the 16-bit modes vs the fp32 mode (run on a GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, "tests/golden"); sys.path.insert(0, ".")
import synth
from test_gpu_models import make_pcnet, CAM_HW, PRJ_HW, dev
from spaa_b200 import models, ops

torch.backends.cudnn.allow_tf32 = False
P = synth.pcnet_params(31, CAM_HW)
prj0 = synth.textured(32, "pc.prj", (2, 3, *PRJ_HW)).to(dev())
scene = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).expand(2, -1, -1, -1).to(dev())
cot = None
rec = {}
orig_b, orig_f = ops.conv_backward_data, ops.conv_forward


def run(precision, tc=True):
    global cot
    m = models.set_precision(make_pcnet(P, CAM_HW), precision)
    log = []

    def wrap_b(spec, dy, w, in_hw, **kw):
        out = orig_b(spec, dy, w, in_hw, **kw)
        log.append(("bwd %s %d->%d k%d" % (spec.kind, spec.cin, spec.cout, spec.k), out.float().clone()))
        if kw.get("out2") is not None:
            log.append(("bwd-out2", kw["out2"].float().clone()))
        return out

    def wrap_f(spec, x, w, b, **kw):
        out = orig_f(spec, x, w, b, **kw)
        log.append(("fwd %s %d->%d k%d" % (spec.kind, spec.cin, spec.cout, spec.k), out.float().clone()))
        return out
    ops.conv_backward_data, ops.conv_forward = wrap_b, wrap_f
    models.ops.conv_backward_data, models.ops.conv_forward = wrap_b, wrap_f
    ops.TC_ENABLED = tc
    prj = prj0.clone().requires_grad_(True)
    y = m(prj, scene)
    if cot is None:
        cot = synth.randn(34, "pc.cot", y.shape).to(dev())
    (y * cot).sum().backward()
    ops.TC_ENABLED = True
    ops.conv_backward_data, ops.conv_forward = orig_b, orig_f
    log.append(("prj.grad", prj.grad.clone()))
    return log


ref = run("fp32")
for prec, tc in (("bf16", True), ("bf16", False), ("fp16", True)):
    try:
        got = run(prec, tc)
    except Exception as e:
        print(prec, tc, "failed:", e)
        continue
    print(f"---- {prec} tc={tc}")
    for (n, a), (n2, b) in zip(got, ref):
        assert n == n2, (n, n2)
        rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        frac_zero_mismatch = ((a == 0) != (b == 0)).float().mean().item()
        print(f"{n:28s} rel-F err {rel:9.2e}   zero-pattern mismatch {frac_zero_mismatch:8.2e}   |ref| {b.abs().mean().item():.3e}")
