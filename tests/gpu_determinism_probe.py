"""Run-to-run determinism of the drop-in model (diagnostic): eval mode has no atomics in the forward path and must be
bit-identical; train mode differs only through the fp32 atomics of the BatchNorm sums."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200")]
import torch
from b200caps import engine
from models.capsules_ucf101 import CapsNet
from oracle import restate

sd = restate.make_state_dict(24, seed=0)
m = CapsNet(pt_path=None); m.load_state_dict(sd); m = m.cuda()
b = restate.synthetic_batch(1, 1, seed=47)
masks = restate.make_drop_masks(2, seed=3, count=4)
data, action, labels = b["data"].cuda(), b["action"].cuda(), b["labels"].cuda()

def l2(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

def run(train):
    m.train(train)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    it = iter(masks)
    engine.STATE.dropout_source = lambda n, c, dev: next(it).reshape(n, c)
    with torch.no_grad():
        x_cl, c56, c112, drop2 = m._encode(data)
        caps, rout = m._capsules(x_cl)
        out, act, feat = m._decode(rout, x_cl, c56, c112, drop2, action, labels, 1 if train else 0, 11 if train else 0)
    engine.STATE.dropout_source = None
    m.load_state_dict(sd0)
    return dict(c112=c112.float(), c56=c56.float(), x=x_cl.float(), caps=caps, rout_mu=rout[..., :384], rout_a=rout[..., 384:], out=out)

for train in (False, True):
    a, bb = run(train), run(train)
    print("train" if train else "eval ", {k: f"{l2(a[k], bb[k]):.1e}" for k in a},
          "per-clip logits L2:", [f"{l2(a['out'][i], bb['out'][i]):.1e}" for i in range(2)])
