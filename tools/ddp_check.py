"""2-GPU check of data-parallel training (launch with torchrun): after a few steps (eager + CUDA-graph replay, gradient exchange overlapped with
the backward tail or not) every rank must hold bit-identical parameters, and both exchange schedules must follow the same loss trajectory."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, torch.nn as nn, torch.distributed as dist
import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
from spaa_b200 import models, train_network as tn
N, hw, phw = 12, (48, 64), (64, 64)
P = synth.pcnet_params(81, hw)
prj = synth.textured(82, "ddp.prj", (N, 3, *phw)).to(dev); scene = synth.textured(83, "ddp.scene", (1, 3, *hw)).to(dev); cam = synth.textured(84, "ddp.cam", (N, 3, *hw)).to(dev)
res = {}
for overlap in (True, False):
    os.environ.pop("SPAA_NO_GRAD_OVERLAP", None)
    if not overlap:
        os.environ["SPAA_NO_GRAD_OVERLAP"] = "1"
    m = models.PCNet(P["mask"], nn.DataParallel(models.WarpingNet(out_size=hw)), nn.DataParallel(models.ShadingNetSPAA()))
    m.load_state_dict(P, strict=True)
    m = models.set_precision(m.to(dev), "bf16")
    cfg = tn.AttrDict(device=str(dev), data_root=None, model_name="PCNet", num_train=N, batch_size=8, max_iters=8, lr=1e-3, lr_drop_ratio=0.2, lr_drop_rate=800,
                      l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, iter_offset=401, save_checkpoint=False, dp_mode="global")
    random.seed(5)
    tn.train_pcnet(m, dict(cam_scene=scene, cam_train=cam, prj_train=prj, mask=P["mask"]), None, cfg, verbose=False)
    flat = torch.cat([p.detach().flatten() for p in m.parameters()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    same = all(torch.equal(both[0], b) for b in both[1:])
    res[overlap] = (cfg["loss_history"][:, 0].cpu(), same)
    if rank == 0:
        print(f"overlap={overlap}: ranks hold identical parameters: {same}; losses {[round(v, 5) for v in res[overlap][0].tolist()]}")
if rank == 0:
    d = (res[True][0] - res[False][0]).abs()
    print("max |loss(overlap) - loss(no overlap)| over the first 4 steps:", d[:4].max().item(), " all steps:", d.max().item())
    assert res[True][1] and res[False][1] and d[:4].max().item() < 5e-3
    print("ddp check ok")
dist.destroy_process_group()
