"""Parameter gradients of one PCNet training step: 16-bit tensor-core modes vs the exact fp32 mode (same batch, same weights)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import synth
from spaa_b200 import models, train_network as tn
dev = torch.device("cuda:0")
CAM, PRJ = (240, 320), (256, 256)
B = 4
g = torch.Generator(device=dev).manual_seed(7)
prj = torch.rand(B, 3, *PRJ, device=dev, generator=g)
cam = torch.rand(B, 3, *CAM, device=dev, generator=g)
scene = synth.textured(0, "gp.scene", (1, 3, *CAM)).to(dev).expand(B, -1, -1, -1)
P = synth.pcnet_params(300, CAM)


def grads(prec):
    m = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=CAM)), torch.nn.DataParallel(models.ShadingNetSPAA()))
    m.load_state_dict(P, strict=True)
    m = models.set_precision(m.to(dev), prec)
    out = m(prj, scene)
    loss, _ = tn.compute_loss(out, cam, "l1+ssim")
    loss.backward()
    return loss.item(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


l32, g32 = grads("fp32")
for prec in sys.argv[1:] or ["bf16"]:
    l16, g16 = grads(prec)
    print(f"loss fp32 {l32:.6f}  {prec} {l16:.6f}")
    for n in g32:
        a, b = g32[n].double(), g16[n].double()
        rel = ((a - b).norm() / (a.norm() + 1e-30)).item()
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
        print(f"{n:45s} |g| {a.norm().item():.3e}  rel err {rel:.3e}  cos {cos:.4f}  ratio {b.norm().item() / (a.norm().item() + 1e-30):.3f}")
