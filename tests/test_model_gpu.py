"""-m gpu: the drop-in CapsNet (b200caps kernels) against the oracle restatement on identical deterministic
weights and synthetic clips: forward (train / eval), parameter gradients, golden fixtures of the reference."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def setup():
    from b200caps import engine
    from models.capsules_ucf101 import CapsNet
    from oracle import restate
    torch.set_num_threads(os.cpu_count())
    sd = restate.make_state_dict(24, seed=0)
    model = CapsNet(pt_path=None)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    batch = restate.synthetic_batch(1, 1, seed=47)
    masks = restate.make_drop_masks(2, seed=3, count=4)
    return dict(model=model, sd=sd, batch=batch, masks=masks, engine=engine, restate=restate)


def _inject(engine, masks):
    it = iter(masks)
    engine.STATE.dropout_source = lambda n, c, dev: next(it).reshape(n, c)


def test_eval_forward_matches_reference_golden(setup):
    s = setup
    model, b = s["model"], s["batch"]
    gold = json.load(open(os.path.join(GOLD, "capsnet_fwd_b2.json")))["eval"]
    model.eval()
    with torch.no_grad():
        out, act, feat = model(b["data"].cuda(), b["action"].cuda(), b["labels"].cuda(), 0, 0)
    act_ref = torch.tensor(gold["act"])
    assert rel(act, act_ref) < 2e-2
    assert act.argmax(1).tolist() == gold["argmax"] or float((act_ref.topk(2).values[:, 0] - act_ref.topk(2).values[:, 1]).min()) < 2e-2 * float(act_ref.abs().max())
    f = out.double().cpu().reshape(-1)
    vals = f[torch.tensor(gold["logits"]["idx"])]
    ref = torch.tensor(gold["logits"]["vals"], dtype=torch.float64)
    assert float((vals - ref).abs().max()) / gold["logits"]["maxabs"] < 2e-2


def test_train_forward_and_grads_match_oracle(setup):
    """bf16 mode: logits / activations / per-tensor gradients within 2e-2 (max-abs, normalised) of the fp64 oracle;
    thresholded masks and argmax identical outside the stated margin."""
    s = setup
    model, sd, b, masks, engine, restate = s["model"], s["sd"], s["batch"], s["masks"], s["engine"], s["restate"]
    model.train()
    for p in model.parameters():
        p.grad = None
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    _inject(engine, masks[:2])
    try:
        out, act, feat = model(b["data"].cuda(), b["action"].cuda(), b["labels"].cuda(), 1, 11)
    finally:
        engine.STATE.dropout_source = None
    # oracle fp64 with autograd
    sd64 = {k: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
            for k, v in sd.items()}
    bn = restate.BNState(True)
    o_ref, a_ref, f_ref = restate.capsnet_forward(sd64, b["data"].double(), b["action"], b["labels"], 1, 11, True,
                                                  [m.double() for m in masks[:2]], bn)
    e_logits, e_act, e_feat = rel(out, o_ref.detach()), rel(act, a_ref.detach()), rel(feat, f_ref.detach())
    print(f"train fwd: logits {e_logits:.2e} act {e_act:.2e} feat {e_feat:.2e}")
    assert e_logits < 2e-2 and e_act < 2e-2 and e_feat < 5e-2
    # thresholded masks: bit-exact where the oracle margin exceeds the tolerance
    tol = 2e-2 * float(o_ref.abs().max())
    safe = o_ref.detach().abs() > tol
    agree = ((out.cpu() > 0) == (o_ref.detach() > 0)) | ~safe
    assert bool(agree.all()), f"{int((~agree).sum())} mask flips outside the margin; excluded {int((~safe).sum())}"
    # BN running statistics
    new_sd = model.state_dict()
    worst = 0.0
    for p, (rm, rv) in bn.updates.items():
        worst = max(worst, rel(new_sd[p + ".bn.running_mean"], rm), rel(new_sd[p + ".bn.running_var"], rv))
    assert worst < 2e-2, worst
    # gradients of a scalar that touches all three outputs
    g = torch.Generator().manual_seed(9)
    w_o = torch.randn(out.shape, generator=g) / out.numel() ** 0.5
    w_a = torch.randn(act.shape, generator=g)
    w_f = torch.randn(feat.shape, generator=g) / 400
    loss = (out * w_o.cuda()).sum() + (act * w_a.cuda()).sum() + (feat * w_f.cuda()).sum()
    loss.backward()
    loss_ref = (o_ref * w_o.double()).sum() + (a_ref * w_a.double()).sum() + (f_ref * w_f.double()).sum()
    names = [k for k, v in sd64.items() if v.requires_grad]
    grads_ref = torch.autograd.grad(loss_ref, [sd64[k] for k in names])
    gp = dict(model.named_parameters())
    bad = []
    worst = 0.0
    for k, gr in zip(names, grads_ref):
        e = rel(gp[k].grad, gr)
        worst = max(worst, e)
        if e > 5e-2:
            bad.append((k, e))
    print(f"worst per-tensor gradient error {worst:.2e}")
    assert not bad, bad[:10]
    model.load_state_dict(sd0)
