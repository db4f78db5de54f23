"""Run each tensor-core backward-weight case in its own process (a trap poisons the CUDA context)."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("conv", 128, 256, 3, 1, 1, 0, 20, 24), ("conv", 256, 128, 3, 1, 1, 0, 17, 9), ("conv", 64, 128, 3, 1, 1, 0, 16, 32),
         ("conv", 32, 64, 1, 1, 0, 0, 24, 32), ("conv", 32, 64, 3, 2, 1, 0, 26, 38), ("conv", 64, 64, 3, 2, 1, 0, 16, 32),
         ("convT", 128, 64, 3, 2, 1, 1, 7, 11), ("convT", 64, 32, 2, 2, 0, 0, 12, 16), ("conv", 32, 16, 3, 1, 1, 0, 20, 36),
         ("conv", 16, 32, 3, 2, 1, 0, 20, 36), ("conv", 64, 64, 1, 1, 0, 0, 16, 8), ("conv", 128, 128, 1, 1, 0, 0, 16, 8)]
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests", "golden"))
import torch, torch.nn.functional as F, synth
from spaa_b200 import ops
kind, cin, cout, k, stride, pad, outpad, H, W = %r
adt = torch.float16 if %r == "mix" else torch.bfloat16
B = 3
spec = ops.ConvSpec(kind, cin, cout, k, stride, pad, outpad)
x = synth.randn(31, "wg.x", (B, cin, H, W)).to(adt)
w = torch.zeros(spec.weight_shape(), dtype=torch.double, requires_grad=True)
pre = F.conv2d(x.double(), w, None, stride, pad) if kind == "conv" else F.conv_transpose2d(x.double(), w, None, stride, pad, outpad)
dy = synth.randn(32, "wg.dy", pre.shape).to(torch.bfloat16)
gw, = torch.autograd.grad((pre * dy.double()).sum(), w)
cl = lambda t: t.to("cuda:0").contiguous(memory_format=torch.channels_last)
dw = torch.zeros(spec.weight_shape(), device="cuda:0")
ops.conv_backward_weight(spec, cl(x), cl(dy), dw, None)
torch.cuda.synchronize()
err = (dw.cpu().double() - gw).abs().max().item()
print("max err %%.3e  scale %%.3e" %% (err, gw.abs().max().item()))
'''
for mix in ("bf16", "mix"):
    for c in CASES:
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT, c, mix)], capture_output=True, text=True, timeout=120)
        out = (r.stdout.strip().splitlines() or [""])[-1]
        err = [l for l in r.stderr.splitlines() if "Error" in l or "error" in l][-1:] if r.returncode else []
        print(mix, c, "->", out if r.returncode == 0 else f"FAIL {err}", flush=True)
