import sys, torch, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'tests/golden'); sys.path.insert(0,'.')
import synth
from oracle import spaa_oracle as O
from test_gpu_models import _spaa_setup, TinyClf, LABELS, SETUP, CAM_HW, PRJ_HW, dev
from spaa_b200 import projector_based_attack as pba
torch.backends.cudnn.allow_tf32=False
P, m, scene = _spaa_setup()
targets=[int(v) for v in np.load('tests/golden/spaa.npz')['targets']]
tiny=synth.TinyClassifier(1)
otr=[]
O.spaa_attack(lambda x, s: O.pcnet(P, x, s, CAM_HW), lambda im: O.classify(tiny, im, (24, 24), (20, 20)), targets, True, scene, 2.0, "camdE_caml2", prj_hw=PRJ_HW, iters=10, trace=otr)
tr=[]
pba.spaa(m, TinyClf(1), LABELS, targets, True, scene, 2.0, "camdE_caml2", dev(), SETUP, iters=10, trace=tr, forced_prj=[t['prj_in'].to(dev()) for t in otr])
for i,(a,o) in enumerate(zip(tr,otr)):
    sr=o['prj_out']-o['prj_in']; sg=(a['prj_out']-a['prj_in']).cpu()
    e=(sg-sr).abs()
    bad=(e>2e-5+1e-3*sr.abs())
    print(i,'use_col',o['use_col'].int().tolist(),'nbad',int(bad.sum()),'per-sample nbad',bad.flatten(1).sum(1).tolist(),'max',e.max().item())
    if bad.any():
        idx=bad.nonzero()[:8].tolist(); print('   ',idx, [ (sg[tuple(j)].item(), sr[tuple(j)].item()) for j in idx[:4]])
        # outside-range pixels?
        pin=o['prj_in']
        print('    prj_in at bad:', [pin[tuple(j)].item() for j in idx[:8]])
