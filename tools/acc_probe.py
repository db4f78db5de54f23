"""How do the B200 tensor cores round the fp32 accumulation of bf16 / tf32 products?  (Decides whether a split-precision "fp32-accurate" convolution on
tcgen05 can meet 1e-5.)  Products are exact in fp32; the running sum needs more than 24 bits, so every accumulation step rounds:
round-to-nearest errors average out (~sqrt(n) ulp), truncation biases the sum low by ~n/2 ulp."""
import torch
dev = torch.device("cuda:0")
torch.manual_seed(0)
for K in (1024, 4096, 16384):
    # positive products p = (1 + 2^-7 * r)(1 + 2^-7 * s), r, s in {0..127}: bf16-exact factors, fp32-exact products
    r = torch.randint(0, 128, (128, K), device=dev).float()
    s = torch.randint(0, 128, (K, 128), device=dev).float()
    a, b = (1 + r / 128), (1 + s / 128)
    exact = (a.double() @ b.double())
    for name, fn in (("bf16 x bf16 -> fp32 (cuBLAS)", lambda: torch.mm(a.bfloat16(), b.bfloat16(), out_dtype=torch.float32)),
                     ("tf32 (allow_tf32)", None), ("fp32 SIMT/exact", None)):
        if name.startswith("tf32"):
            torch.backends.cuda.matmul.allow_tf32 = True
            got = a @ b
            torch.backends.cuda.matmul.allow_tf32 = False
        elif name.startswith("fp32"):
            got = a @ b
        else:
            try:
                got = fn()
            except Exception as e:
                print(name, "unavailable:", e); continue
        err = (got.double() - exact)
        ulp = torch.finfo(torch.float32).eps * exact.abs()
        print(f"K={K:6d} {name:32s} mean err {err.mean().item():+.3e} ({(err / ulp).mean().item():+.2f} ulp)  rms {(err / ulp).pow(2).mean().sqrt().item():.2f} ulp  max {(err/ulp).abs().max().item():.1f} ulp   rel max {(err.abs() / exact.abs()).max().item():.2e}")
