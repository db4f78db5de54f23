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


for k, s, p, cout, hw in ((7, 2, 3, 64, 224),):
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

# space-to-depth form of the 7x7 s2 p3 stem: 4x4 s1 p0 convolution over the 2x2-folded, explicitly padded input [B, 12 (or 16), 115, 115]
for C in (12, 16):
    for k, hw in ((4, 115), (5, 116)):
        torch.manual_seed(0)
        conv = torch.nn.Conv2d(C, 64, k, 1, 0).to(dev).to(memory_format=torch.channels_last)
        for q in conv.parameters():
            q.requires_grad = False
        x = torch.randn(B, C, hw, hw, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = conv(x)
        dy = torch.randn_like(y)
        fwd = timeit(lambda: conv(x))
        both = timeit(lambda: torch.autograd.grad(conv(x), x, dy))
        print(f"s2d conv {k}x{k} s1 {C}->64 @{hw} -> {tuple(y.shape[2:])}: fwd {fwd:7.1f} us   fwd+dgrad {both:7.1f} us", flush=True)
for bench in (True,):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(12, 64, 4, 1, 0).to(dev).to(memory_format=torch.channels_last)
    for q in conv.parameters():
        q.requires_grad = False
    x = torch.randn(B, 12, 115, 115, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = torch.randn_like(conv(x))
    print(f"s2d conv 4x4 12->64 cudnn.benchmark={bench}: fwd {timeit(lambda: conv(x)):7.1f} us   fwd+dgrad {timeit(lambda: torch.autograd.grad(conv(x), x, dy)):7.1f} us", flush=True)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3).to(dev).to(memory_format=torch.channels_last)
    for q in conv.parameters():
        q.requires_grad = False
    x = torch.randn(B, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = torch.randn_like(conv(x))
    print(f"7x7 s2 3->64 cudnn.benchmark={bench}: fwd {timeit(lambda: conv(x)):7.1f} us   fwd+dgrad {timeit(lambda: torch.autograd.grad(conv(x), x, dy)):7.1f} us", flush=True)
