import sys, os, random
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden"); sys.path.insert(0, "/root/repo/tests")
import torch, torch.nn as nn, synth
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from spaa_b200 import models, train_network as tn
N, hw, phw = 6, (48, 64), (64, 64)
P = synth.pcnet_params(81, hw)
prj_train = synth.textured(82, "trg.prj", (N, 3, *phw)); scene = synth.textured(83, "trg.scene", (1, 3, *hw)); cam_train = synth.textured(84, "trg.cam", (N, 3, *hw))
def run(graph, prec="fp32"):
    m = models.PCNet(P["mask"], nn.DataParallel(models.WarpingNet(out_size=hw)), nn.DataParallel(models.ShadingNetSPAA()))
    m.load_state_dict(P, strict=True)
    m = nn.DataParallel(models.set_precision(m.to("cuda:0"), prec), device_ids=[0])
    cfg = tn.AttrDict(device="cuda:0", data_root=None, model_name="PCNet", num_train=N, batch_size=4, max_iters=8, lr=1e-3, lr_drop_ratio=0.2,
                      lr_drop_rate=800, l2_reg=1e-4, plot_on=False, valid_rate=10 ** 9, iter_offset=0, save_checkpoint=False, graph=graph)
    random.seed(5)
    tn.train_pcnet(m, dict(cam_scene=scene, cam_train=cam_train, prj_train=prj_train, mask=P["mask"]), None, cfg, verbose=False)
    return cfg["loss_history"][:, 0].cpu().double()
e1, e2, g1, g2 = run(False), run(False), run(True), run(True)
print("eager-eager", (e1 - e2).abs().tolist())
print("graph-graph", (g1 - g2).abs().tolist())
print("graph-eager", (g1 - e1).abs().tolist())
