"""Diagnostic: parameter gradients of CompenNet++ / PCNet at the toy test size in 'fp32' (CUDA cores) and 'bf16x3' (split-precision tensor cores) against
float64 autograd of the oracle; and the number of ReLU masks of the surface branch that differ between the two modes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, torch.nn as nn
import synth
from oracle import spaa_oracle as O
from spaa_b200 import models
from spaa_b200.models import _Stack

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
CAM_HW, PRJ_HW = (24, 32), (32, 32)
C = synth.compennet_pp_params(37)
cam = synth.textured(38, "cpp.cam", (2, 3, *CAM_HW))
scene = synth.textured(33, "pc.s", (1, 3, *CAM_HW)).expand(2, -1, -1, -1)
cot = synth.randn(39, "cpp.cot", (2, 3, *PRJ_HW))
P64 = {k: v.double().requires_grad_(v.dtype.is_floating_point and "ctrl" not in k) for k, v in C.items()}
y64 = O.compennet_pp(P64, cam.double(), scene.double(), PRJ_HW)
names = [k for k, v in P64.items() if v.requires_grad]
g64 = dict(zip(names, torch.autograd.grad((y64 * cot.double()).sum(), [P64[k] for k in names], allow_unused=True)))
res = {}
acts = {}
for prec in ("fp32", "bf16x3"):
    m = models.CompenNetPlusplus(nn.DataParallel(models.WarpingNet(out_size=PRJ_HW)), nn.DataParallel(models.CompenNet()))
    m.load_state_dict(C, strict=True)
    m = models.set_precision(m.to(dev), prec)
    # keep the saved activations of the stack
    keep = {}
    orig = _Stack.forward
    def fwd(*a, _o=orig, **k):
        out, S = _o(*a, **k)
        keep.update({n: S[n] for n in ("r1s", "r2s", "r3s", "r4s", "x1", "x2", "x3", "x4", "x5", "x6", "x7")})
        return out, S
    _Stack.forward = staticmethod(fwd)
    y = m(cam.to(dev), scene.to(dev))
    _Stack.forward = staticmethod(orig)
    (y * cot.to(dev)).sum().backward()
    res[prec] = {n: p.grad.detach().double().cpu() for n, p in m.named_parameters()}
    def logical(t):
        if prec == "bf16x3":
            c = t.shape[1] // 3
            return (t[:, :c].double() + t[:, c:2 * c].double() + t[:, 2 * c:].double()).cpu()
        return t.double().cpu()
    acts[prec] = {n: logical(t) for n, t in keep.items()}
    print(prec, "output max err vs float64:", (y.detach().double().cpu() - y64.detach()).abs().max().item())
print(f"{'parameter':40s} {'fp32 rel-max err':>18s} {'bf16x3 rel-max err':>18s}")
for n in res["fp32"]:
    ref = g64.get(n)
    if ref is None:
        continue
    sc = ref.abs().max().item() + 1e-30
    print(f"{n:40s} {(res['fp32'][n] - ref).abs().max().item() / sc:18.2e} {(res['bf16x3'][n] - ref).abs().max().item() / sc:18.2e}")
for n in acts["fp32"]:
    a, b = acts["fp32"][n], acts["bf16x3"][n]
    print(f"{n}: sign mismatches {(a > 0).ne(b > 0).sum().item()} of {a.numel()}, max abs diff {(a - b).abs().max().item():.2e}, smallest positive {min(a[a > 0].min().item(), b[b > 0].min().item()):.2e}")
