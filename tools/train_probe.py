"""Per-kernel time of one PCNet training step (B=24) via torch.profiler.  usage: python tools/train_probe.py [fp32|fp16|bf16]"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import bench, synth
from spaa_b200 import models, train_network as tn
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
N = 48
prj = torch.rand(N, 3, *bench.PRJ_HW, device=dev, generator=g)
cam = torch.rand(N, 3, *bench.CAM_HW, device=dev, generator=g)
scene = synth.textured(0, "bench.train.scene", (1, 3, *bench.CAM_HW)).to(dev)
P = synth.pcnet_params(300, bench.CAM_HW)
model = models.PCNet(P["mask"], torch.nn.DataParallel(models.WarpingNet(out_size=bench.CAM_HW)), torch.nn.DataParallel(models.ShadingNetSPAA()))
model.load_state_dict(P, strict=True)
model = models.set_precision(model.to(dev), prec)


def run(n, offset=401):
    cfg = tn.AttrDict(device=str(dev), data_root=None, model_name="PCNet", num_train=N, batch_size=24, max_iters=n, lr=1e-3, lr_drop_ratio=0.2,
                      lr_drop_rate=800, l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, dp_mode="weak", iter_offset=offset, save_checkpoint=False)
    random.seed(1)
    tn.train_pcnet(model, dict(cam_scene=scene, cam_train=cam, prj_train=prj, mask=P["mask"]), None, cfg, verbose=False)
    return cfg


run(2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); cfg = run(5); e1.record(); torch.cuda.synchronize()
print(f"precision {prec}: {e0.elapsed_time(e1) / 5:.2f} ms/step, losses {cfg['loss_history'][:, 0].tolist()}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(2)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda k: -k.device_time_total)[:28]
tot = sum(k.device_time_total for k in prof.key_averages())
print(f"total device time for 2 steps: {tot / 1e3:.2f} ms")
for k in rows:
    print(f"{k.device_time_total / 2e3:9.3f} ms/step  x{k.count // 2:4d}  {k.key[:110]}")
wg = [e for e in prof.events() if "conv_wgrad_tc_kernel" in e.name and e.device_time_total > 0]
half = wg[len(wg) // 2:]
print("conv_wgrad_tc_kernel launches of the last step (us):", [round(e.device_time_total, 1) for e in half])
