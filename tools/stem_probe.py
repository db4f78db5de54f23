"""How long does the external classifier's first cuDNN convolution (resnet18 conv1: 7x7 s2, 3 -> 64, B = 32, 224x224, channels_last, TF32) take forward +
input-gradient as a function of the INPUT channel padding (3, 4, 8 zero-padded channels: same function, different cuDNN kernels)?  Diagnostic only."""
import sys
import torch

dev = torch.device("cuda:0")
B = 32


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for k, s, p, cout, hw in ((7, 2, 3, 64, 224), (3, 1, 1, 64, 224), (3, 2, 0, 32, 299)):
    for C in (3, 4, 8):
        torch.manual_seed(0)
        conv = torch.nn.Conv2d(C, cout, k, s, p).to(dev).to(memory_format=torch.channels_last)
        for q in conv.parameters():
            q.requires_grad = False
        x = torch.randn(B, C, hw, hw, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = conv(x)
        dy = torch.randn_like(y)
        fwd = timeit(lambda: conv(x))
        both = timeit(lambda: torch.autograd.grad(conv(x), x, dy))
        print(f"conv {k}x{k} s{s} {C}->{cout} @{hw}: fwd {fwd:7.1f} us   fwd+dgrad {both:7.1f} us", flush=True)
