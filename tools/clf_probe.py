"""How fast can the EXTERNAL classifier leg (torchvision resnet18, cuDNN) run as-is?  Times fwd + bwd-to-input at B=32 under
memory-format / cuDNN-autotune / CUDA-graph variants.  Diagnostic only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchvision import models

dev = torch.device("cuda:0")
B = 32
name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
sz = 299 if name == "inception_v3" else 224


def make(cl):
    torch.manual_seed(0)
    kw = dict(weights=None)
    if name == "inception_v3":
        kw.update(init_weights=False, transform_input=True, aux_logits=True)
    net = getattr(models, name)(**kw).to(dev).eval()
    for p in net.parameters():
        p.requires_grad = False
    if cl:
        net = net.to(memory_format=torch.channels_last)
    return net


def step(net, x, tgt):
    leaf = x.detach().requires_grad_(True)
    out = net(leaf)
    out = out.logits if hasattr(out, "logits") else out
    loss = -out.gather(1, tgt.view(-1, 1)).sum()
    g, = torch.autograd.grad(loss, leaf)
    return out.detach(), g


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tgt = torch.arange(B, device=dev) * 7 % 1000
for cl in (False, True):
    for bench in (False, True):
        for tf32 in (True, False):
            torch.backends.cudnn.benchmark = bench
            torch.backends.cudnn.allow_tf32 = tf32
            net = make(cl)
            x = torch.rand(B, 3, sz, sz, device=dev)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            ms = timeit(lambda: step(net, x, tgt))
            # CUDA graph of the same
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    step(net, x, tgt)
            torch.cuda.current_stream().wait_stream(s)
            try:
                with torch.cuda.graph(g):
                    o, gr = step(net, x, tgt)
                msg = timeit(g.replay)
            except Exception as e:
                msg = float("nan")
            print(f"{name} channels_last={cl!s:5} cudnn.benchmark={bench!s:5} tf32={tf32!s:5}: eager {ms:.3f} ms   graph {msg:.3f} ms", flush=True)
