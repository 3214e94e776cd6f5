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


@pytest.fixture
def precision(request):
    """Activation / GEMM-operand precision mode of the CUDA path for one test (restored to bf16 afterwards).
    tf32 = fp32 activations + tcgen05 kind::tf32: the mode the north_star 1e-3 bar applies to."""
    from b200caps import plans
    plans.set_precision(request.param)
    yield request.param
    plans.set_precision("bf16")


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


def _cl2ncdhw(t):
    return t.detach().float().permute(0, 4, 1, 2, 3).double().cpu()


@pytest.mark.parametrize("precision", ["bf16", "tf32"], indirect=True)
def test_train_mode_segment_parity(setup, precision):
    """Train mode (batch-stat BN, injected Dropout3d masks), both precision modes, against the fp64 oracle.
    Every bound below is written T(bf16 bound, tf32 bound); the tf32 bounds are the north_star fp32-mode bar (1e-3) for
    the per-segment forwards and a small multiple of it for gradients (noise-like sums, see below).

    The EM routing is chaotically sensitive at random init: bf16 rounding upstream moves the logits by O(1) in the
    REFERENCE ITSELF (oracle with emulated bf16 roundings vs exact: logits 0.59, feat 0.36).  Parity
    is therefore asserted per segment with the routing teacher-forced, within the bf16 tolerance of north_star
    (2e-2 max-abs normalised); the end-to-end deviation is measured and printed next to the oracle's own."""
    s = setup
    model, sd, b, masks, engine, restate = s["model"], s["sd"], s["batch"], s["masks"], s["engine"], s["restate"]
    tf32 = precision == "tf32"
    T = lambda bf, tf: tf if tf32 else bf
    emulate = restate.emulate_tf32 if tf32 else restate.emulate_bf16
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    for p in model.parameters():
        p.grad = None
    sd64 = {k: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
            for k, v in sd.items()}
    m64 = [m.double() for m in masks]
    img = b["data"].cuda()

    # ---- segment 1: encoder ------------------------------------------------------------------------
    _inject(engine, masks[:2])
    try:
        x_cl, c56, c112, drop2 = model._encode(img)
    finally:
        engine.STATE.dropout_source = None
    # bf16-mode yardstick for the encoder = the oracle with the SAME bf16 rounding points (GEMM operands, stored
    # activations).  Train-mode BatchNorm over a 2-clip batch amplifies bf16 rounding ~30x by Mixed_4f in the
    # reference itself (exact fp64 vs bf16-rounded reference: x 1.9e-1); that deviation is printed, not asserted.
    bn = restate.BNState(True)
    with emulate():
        x_ref, c56_ref, c112_ref = restate.encode(sd64, b["data"].double(), m64[0], bn)
    with torch.no_grad():
        x_ex, c56_ex, c112_ex = restate.encode({k: v.detach() for k, v in sd64.items()}, b["data"].double(), m64[0],
                                               restate.BNState(True))
    e = dict(x=rel(_cl2ncdhw(x_cl)[:, :, 0], x_ref.detach()), c56=rel(_cl2ncdhw(c56), c56_ref.detach()),
             c112=rel(_cl2ncdhw(c112), c112_ref.detach()))
    e_ex = dict(x=rel(_cl2ncdhw(x_cl)[:, :, 0], x_ex), c56=rel(_cl2ncdhw(c56), c56_ex), c112=rel(_cl2ncdhw(c112), c112_ex))
    o_ex = dict(x=rel(x_ref.detach(), x_ex), c56=rel(c56_ref.detach(), c56_ex), c112=rel(c112_ref.detach(), c112_ex))
    print(f"[{precision}] encoder vs oracle(same roundings):", e)
    print(f"[{precision}] encoder vs exact fp64 oracle     :", e_ex)
    print(f"[{precision}] oracle(same roundings) vs exact  :", o_ex)
    # shallow taps: within the bf16 tolerance of the exact oracle.  The deep tap (17 conv+BN layers) is compared
    # with the rounding-emulated oracle; accumulation-order differences flip bf16 roundings and train-mode BN over a
    # 2-clip batch amplifies them, so its max-norm bound is loose and an L2 bound is added.
    # vs the oracle with the same rounding points: the mode's bar (tf32: 1.5 x, both sides carry rounding noise); vs the
    # exact oracle: the bar or 1.5 x what the reference's own mathematics moves by under that rounding, whichever is larger
    bar = T(2e-2, 1e-3)
    for k in ("c112", "c56"):
        assert e[k] < T(2e-2, 1.5e-3), (k, e)
        assert e_ex[k] < max(bar, 1.5 * o_ex[k]), (k, e_ex, o_ex)
    xa, xb = _cl2ncdhw(x_cl)[:, :, 0], x_ref.detach()
    l2 = float((xa - xb).norm() / xb.norm())
    print(f"encoder deep tap: max-norm {e['x']:.2e}, relative L2 {l2:.2e} (vs oracle with bf16 roundings)")
    assert e["x"] < T(0.2, 2.5e-2) and l2 < T(0.1, 2.5e-2), (e, l2)      # tf32: the oracle's own rounded-vs-exact is 2.9e-2
    new_sd = model.state_dict()
    worst = max(max(rel(new_sd[p + ".bn.running_mean"], rm), rel(new_sd[p + ".bn.running_var"], rv))
                for p, (rm, rv) in bn.updates.items())
    assert worst < T(2e-2, 1e-3), worst
    # encoder gradients through a linear probe of the three taps
    g = torch.Generator().manual_seed(9)
    w1 = torch.randn(x_ref.shape, generator=g, dtype=torch.float64)
    w2 = torch.randn(c56_ref.shape, generator=g, dtype=torch.float64) * 0.2
    w3 = torch.randn(c112_ref.shape, generator=g, dtype=torch.float64) * 0.05
    enc_names = [k for k in sd64 if k.startswith("conv1.") and sd64[k].requires_grad]
    gref = torch.autograd.grad((x_ref * w1).sum() + (c56_ref * w2).sum() + (c112_ref * w3).sum(), [sd64[k] for k in enc_names])
    to_cl_w = lambda w: w.permute(0, 2, 3, 4, 1).float().cuda()
    probe = (x_cl.float() * to_cl_w(w1.unsqueeze(2))).sum() + (c56.float() * to_cl_w(w2)).sum() + (c112.float() * to_cl_w(w3)).sum()
    probe.backward()
    gp = dict(model.named_parameters())
    errs = {k: rel(gp[k].grad, gr) for k, gr in zip(enc_names, gref)}
    worst_k = max(errs, key=errs.get)
    # Reported only: through 17 train-mode BN layers on a 2-clip batch the ORACLE'S OWN gradients move by a median of
    # 75 % (max-abs, normalised) when its GEMM operands are rounded to bf16 (measured); gradient parity of
    # the encoder is asserted module by module in test_encoder_modules_fwd_bwd below.
    print(f"encoder grads (whole trunk, reported): worst {worst_k} {errs[worst_k]:.2e}; median {sorted(errs.values())[len(errs) // 2]:.2e}")
    assert all(torch.isfinite(gp[k].grad).all() for k in enc_names)
    for p in model.parameters():
        p.grad = None

    # ---- segment 2: PrimaryCaps (linear) + routing (teacher-forced on OUR capsules) ---------------------
    x_leaf = x_cl.detach().requires_grad_(True)
    caps, rout = model._capsules(x_leaf)
    x_in = _cl2ncdhw(x_cl)[:, :, 0].requires_grad_(True)
    caps_ref = restate.primary_caps(x_in, sd64)
    e_caps = rel(caps, caps_ref.detach())
    print(f"[{precision}] primary caps (same input): {e_caps:.2e}")
    assert e_caps < T(1e-2, 1e-3)
    wc = torch.randn(caps_ref.shape, generator=g, dtype=torch.float64)
    pc_names = [k for k in sd64 if k.startswith("primary_caps.")]
    gref = torch.autograd.grad((caps_ref * wc).sum(), [sd64[k] for k in pc_names] + [x_in])
    (caps * wc.float().cuda()).sum().backward(retain_graph=True)
    for k, gr in zip(pc_names, gref[:-1]):
        # tf32: the 32-column activation branch carries the smallest gradients (measured 3.1e-3; dz = g a (1 - a) is stored
        # tf32-rounded before the weight-gradient GEMM)
        assert rel(gp[k].grad, gr) < T(3e-2, 5e-3), (k, rel(gp[k].grad, gr))
    assert rel(_cl2ncdhw(x_leaf.grad)[:, :, 0], gref[-1]) < T(2e-2, 2e-3)
    for p in model.parameters():
        p.grad = None
    caps_d = caps.detach().double().cpu()
    mu_tf, a_tf = restate.em_routing(caps_d[..., :512].reshape(800, 32, 16), caps_d[..., 512:].reshape(800, 32),
                                     sd64["conv_caps.weights"][0].detach(), sd64["conv_caps.beta_u"].detach(),
                                     sd64["conv_caps.beta_a"].detach())
    e_mu = rel(rout[..., :384].reshape(800, 24, 16), mu_tf)
    e_a = rel(rout[..., 384:].reshape(800, 24), a_tf)
    print(f"routing (teacher-forced): mu {e_mu:.2e} a {e_a:.2e}")
    assert e_mu < 1e-3 and e_a < 1e-3

    # ---- segment 3: class activations, pose mask, decoder (teacher-forced on OUR routing output) -------
    rout_leaf = rout.detach().requires_grad_(True)
    out, act, feat = model._decode(rout_leaf, x_cl.detach(), c56.detach(), c112.detach(), drop2, b["action"].cuda(),
                                   b["labels"].cuda(), 1, 11)
    rout_in = rout.detach().double().cpu().requires_grad_(True)
    o_ref, a_ref, f_ref = restate.decode(sd64, rout_in, _cl2ncdhw(x_cl)[:, :, 0], _cl2ncdhw(c56), _cl2ncdhw(c112),
                                         b["action"], b["labels"], 1, 11, True, m64[1])
    e = dict(logits=rel(out, o_ref.detach()), act=rel(act, a_ref.detach()), feat=rel(feat, f_ref.detach()))
    print(f"[{precision}] decoder:", e)
    assert e["logits"] < T(2e-2, 1e-3) and e["act"] < 1e-5 and e["feat"] < 1e-6, e
    tol = T(2e-2, 1e-3) * float(o_ref.abs().max())
    safe = o_ref.detach().abs() > tol
    agree = ((out.cpu() > 0) == (o_ref.detach() > 0)) | ~safe
    print(f"[{precision}] thresholded masks: {int((~safe).sum())} of {safe.numel()} pixels inside the margin (excluded)")
    assert bool(agree.all()), f"{int((~agree).sum())} mask flips outside the margin"
    if tf32:      # bit-exact masks everywhere but a < 3 % band around zero (bf16 mode: the 2e-2 margin swallows a third)
        assert int((~safe).sum()) < 0.03 * safe.numel()
    assert act.argmax(1).tolist() == a_ref.argmax(1).tolist()
    w_o = torch.randn(out.shape, generator=g, dtype=torch.float64) / out.numel() ** 0.5
    w_a = torch.randn(act.shape, generator=g, dtype=torch.float64)
    w_f = torch.randn(feat.shape, generator=g, dtype=torch.float64) / 400
    dec_names = [k for k in sd64 if sd64[k].requires_grad and not k.startswith(("conv1.", "primary_caps.", "conv_caps."))]
    gref_exact = torch.autograd.grad((o_ref * w_o).sum() + (a_ref * w_a).sum() + (f_ref * w_f).sum(),
                                     [sd64[k] for k in dec_names] + [rout_in])
    # gradient yardstick = oracle with the same bf16 rounding points: at random init the gradients are noise-like sums,
    # and a fraction f of ReLU masks flipped by bf16 rounding perturbs them by ~sqrt(f) (5-15 %) in the reference itself
    with emulate():
        o_em, a_em, f_em = restate.decode(sd64, rout_in, _cl2ncdhw(x_cl)[:, :, 0], _cl2ncdhw(c56), _cl2ncdhw(c112),
                                          b["action"], b["labels"], 1, 11, True, m64[1])
    gref = torch.autograd.grad((o_em * w_o).sum() + (a_em * w_a).sum() + (f_em * w_f).sum(),
                               [sd64[k] for k in dec_names] + [rout_in])
    ((out * w_o.float().cuda()).sum() + (act * w_a.float().cuda()).sum() + (feat * w_f.float().cuda()).sum()).backward()
    errs = {k: rel(gp[k].grad, gr) for k, gr in zip(dec_names, gref[:-1])}
    errs["rout"] = rel(rout_leaf.grad, gref[-1])
    errs_exact = {k: rel(gp[k].grad, gr) for k, gr in zip(dec_names, gref_exact[:-1])}
    print(f"[{precision}] decoder grads vs oracle(same roundings):", {k: f"{v:.1e}" for k, v in errs.items()})
    print(f"[{precision}] decoder grads vs exact fp64 oracle      :", {k: f"{v:.1e}" for k, v in errs_exact.items()})
    # Bound per tensor: the mode's bar, or -- where the REFERENCE'S OWN gradient moves by more than that when its GEMM
    # operands / stored activations are rounded (own[k] = rounded oracle vs exact oracle: ReLU masks flip, and these
    # gradients are noise-like sums at random init) -- 1.5 x that deviation.  This is the root cause of the round-1
    # "upsample2/3.weight" finding: tests/gpu_wgrad_bisect.py shows the kernels reproduce torch-fp64 on identical
    # operands to 5e-7; the incoming gradient itself differs by ~8e-2 once a few hundred ReLU decisions flip.
    own = {k: rel(gr, ge) for k, gr, ge in zip(dec_names + ["rout"], gref, gref_exact)}
    print(f"[{precision}] reference's own decoder-gradient deviation under the same rounding:", {k: f"{v:.1e}" for k, v in own.items()})
    # The assertion is norm-wise (relative L2 over the tensor): the max-abs figure of a 100 K-element noise-like tensor is
    # set by its single worst element and swings 6e-2 .. 1.5e-1 between runs of the same build (the train-mode encoder
    # sums with fp32 atomics, so the decoder's inputs differ in the last bit from run to run); it is printed above.
    def l2(a, b):
        a, b = a.double().cpu(), b.double().cpu()
        return float((a - b).norm() / (b.norm() + 1e-300))
    ours = [gp[k].grad for k in dec_names] + [rout_leaf.grad]
    for k, g_ours, gr, ge in zip(dec_names + ["rout"], ours, gref, gref_exact):
        e_l2, own_l2 = l2(g_ours, gr), l2(gr, ge)
        assert e_l2 < max(T(3e-2, 1e-3), 1.5 * own_l2), (k, e_l2, own_l2, errs[k], own[k])

    # ---- end to end, reported (not asserted at 2e-2: see docstring) -----------------------------------
    model.load_state_dict(sd0)
    _inject(engine, masks[:2])
    try:
        with torch.no_grad():
            out, act, feat = model(img, b["action"].cuda(), b["labels"].cuda(), 1, 11)
    finally:
        engine.STATE.dropout_source = None
    with torch.no_grad():
        sdd = {k: v.detach() for k, v in sd64.items()}
        o_ex, a_ex, f_ex = restate.capsnet_forward(sdd, b["data"].double(), b["action"], b["labels"], 1, 11, True, m64[:2])
        with emulate():
            o_em, a_em, f_em = restate.capsnet_forward(sdd, b["data"].double(), b["action"], b["labels"], 1, 11, True, m64[:2])
    print("[%s] END-TO-END vs exact fp64 oracle : logits %.2e act %.2e feat %.2e" % (precision, rel(out, o_ex), rel(act, a_ex), rel(feat, f_ex)))
    print("[%s] oracle(same roundings) vs exact : logits %.2e act %.2e feat %.2e" % (precision, rel(o_em, o_ex), rel(a_em, a_ex), rel(f_em, f_ex)))
    assert rel(act, a_ex) < T(5e-2, 1e-2) and torch.isfinite(out).all()
    model.load_state_dict(sd0)


@pytest.mark.parametrize("precision", ["bf16", "tf32"], indirect=True)
@pytest.mark.parametrize("which", ["stem", "conv2b", "conv2c", "Mixed_3b", "Mixed_4f"])
def test_encoder_modules_fwd_bwd(setup, which, precision):
    """Module-level parity (level 2): each encoder building block, train mode (batch statistics),
    forward + all parameter / input gradients against the fp64 oracle on the same bf16-representable input."""
    s = setup
    model, sd, restate = s["model"], s["sd"], s["restate"]
    tf32 = precision == "tf32"
    T = lambda bf, tf: tf if tf32 else bf
    emulate = restate.emulate_tf32 if tf32 else restate.emulate_bf16
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(sum(ord(c) for c in which))
    one, three = (1, 1, 1), (3, 3, 3)
    cfg = {
        "stem": ("conv1.Conv3d_1a_7x7", (2, 3, 8, 64, 64), lambda x, sdd, bn: restate.unit3d(x, sdd, "conv1.Conv3d_1a_7x7", (7, 7, 7), (2, 2, 2), bn)),
        "conv2b": ("conv1.Conv3d_2b_1x1", (2, 64, 4, 28, 28), lambda x, sdd, bn: restate.unit3d(x, sdd, "conv1.Conv3d_2b_1x1", one, one, bn)),
        "conv2c": ("conv1.Conv3d_2c_3x3", (2, 64, 4, 28, 28), lambda x, sdd, bn: restate.unit3d(x, sdd, "conv1.Conv3d_2c_3x3", three, (2, 1, 1), bn)),
        "Mixed_3b": ("conv1.Mixed_3b", (2, 192, 2, 28, 28), lambda x, sdd, bn: restate.inception(x, sdd, "conv1.Mixed_3b", bn)),
        "Mixed_4f": ("conv1.Mixed_4f", (2, 528, 1, 28, 28), lambda x, sdd, bn: restate.inception(x, sdd, "conv1.Mixed_4f", bn)),
    }[which]
    prefix, shape, ref_fn = cfg
    x = torch.relu(torch.randn(shape, generator=g)).bfloat16().float()
    if which == "stem":
        x = torch.rand(shape, generator=g).bfloat16().float()
    mod = model.conv1
    for part in prefix.split(".")[1:]:
        mod = getattr(mod, part)
    for p in mod.parameters():
        p.grad = None
    xg = x.cuda().requires_grad_(which != "stem")
    y = mod(xg)
    sd64 = {k: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
            for k, v in sd.items() if k.startswith(prefix)}
    xr = x.double().requires_grad_(which != "stem")
    # yardstick: the oracle with the same bf16 rounding points (GEMM operands, stored conv output / activation);
    # the exact fp64 oracle is printed next to it (ReLU masks flipped by bf16 rounding make noise-like gradients
    # differ by ~sqrt(flipped fraction) in the reference itself)
    with emulate():
        yr = ref_fn(xr, sd64, restate.BNState(True))
    w = torch.randn(yr.shape, generator=g, dtype=torch.float64)
    names = [k for k, v in sd64.items() if v.requires_grad]
    leaves = [sd64[k] for k in names] + ([xr] if which != "stem" else [])
    gref = torch.autograd.grad((yr * w).sum(), leaves)
    y_ex = ref_fn(xr, sd64, restate.BNState(True))
    gex = torch.autograd.grad((y_ex * w).sum(), leaves)
    e_fwd, e_fwd_ex = rel(y.float(), yr.detach()), rel(y.float(), y_ex.detach())
    (y.float() * w.float().cuda()).sum().backward()
    gp = dict(model.named_parameters())
    errs = {k: rel(gp[k].grad, gr) for k, gr in zip(names, gref)}
    errs_ex = {k: rel(gp[k].grad, gr) for k, gr in zip(names, gex)}
    if which != "stem":
        errs["input"] = rel(xg.grad.float(), gref[-1])
        errs_ex["input"] = rel(xg.grad.float(), gex[-1])
    worst, worst_ex = max(errs, key=errs.get), max(errs_ex, key=errs_ex.get)
    print(f"[{precision}] {which}: fwd {e_fwd:.2e} (exact oracle {e_fwd_ex:.2e}); grads worst {worst} {errs[worst]:.2e} "
          f"(exact oracle: {worst_ex} {errs_ex[worst_ex]:.2e})")
    model.load_state_dict(sd0)
    assert e_fwd < T(2e-2, 2e-3) and e_fwd_ex < T(3e-2, 1e-3), (e_fwd, e_fwd_ex)
    # gradients: the mode's bar, or -- where the REFERENCE'S OWN gradient moves by more than that when its operands are
    # rounded (own[k]: rounded oracle vs exact oracle; BatchNorm backward is a difference of large sums and ReLU masks
    # flip) -- no further from the rounded reference than the rounded reference is from the exact one
    own = {k: rel(gr, ge) for k, gr, ge in zip(names + (["input"] if which != "stem" else []), gref, gex)}
    print(f"   reference's own gradient deviation under the same rounding: worst {max(own.values()):.2e}")
    for k, v in errs.items():
        assert v < max(T(3e-2, 1e-3), own[k]), (k, v, own[k])


def test_jhmdb_variant_eval_and_step():
    """Config 4: the 21-class JHMDB CapsNet (ConvCaps(32, 21), upsample1 ConvT2d(336 -> 64),
    main_jhmdb.py:338,383).  Eval forward (1 clip; the fp64 oracle of one clip takes a few seconds on the host) against
    the oracle restatement with the same name-keyed weights, then one fused --bv training step (finite losses, gradients
    reach the 21-class-specific tensors)."""
    from b200caps.step import StepArgs, TrainStep
    from models.capsules_jhmdb_semi_sup_pa import CapsNet
    from oracle import restate

    sd = restate.make_state_dict(21, seed=0)
    model = CapsNet(pt_path=None)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    b = restate.synthetic_batch(1, 1, seed=47, num_classes=21)
    with torch.no_grad():
        out, act, _ = model(b["data"][:1].cuda(), b["action"][:1].cuda(), b["labels"][:1].cuda(), 0, 0)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        o_ref, a_ref, _ = restate.capsnet_forward(sd64, b["data"][:1].double(), b["action"][:1].double(), b["labels"][:1].double(),
                                                  0, 0, False, None, None, 21)
    assert act.shape == (1, 21)
    assert rel(act, a_ref) < 2e-2
    # eval mode masks the poses with the ARGMAX class (capsules_ucf101.py:473-479): the logits are only comparable when
    # both sides pick the same class, which the oracle's top-2 margin decides (margin rule of SURVEY 8c)
    top2 = a_ref.topk(2).values[0]
    gap = float(top2[0] - top2[1]) / float(a_ref.abs().max())
    same_class = int(act.argmax(1)) == int(a_ref.argmax(1))
    print(f"jhmdb eval: act dev {rel(act, a_ref):.2e}, top-2 gap {gap:.2e}, same class {same_class}, "
          f"logit dev max {rel(out, o_ref):.2e}, rel-L2 {float((out.double().cpu() - o_ref).norm() / o_ref.norm()):.2e}")
    assert same_class or gap < 2e-2
    if same_class:
        assert float((out.double().cpu() - o_ref).norm() / o_ref.norm()) < 2e-2
        assert rel(out, o_ref) < 0.15     # max-abs: single pixels of the 2e-2-sized eval logits, bf16 through 8 layers

    model.train()
    step = TrainStep(model, StepArgs(bv=True, gv=False, n_frames=5, wt_cons=0.1, lr=1e-4))
    r = step(b["data"].cuda(), b["fl_data"].cuda(), b["action"].cuda(), b["seg"].cuda(), b["labels"], epoch=1)
    for k in ("total", "bce", "dice", "cls", "cons"):
        assert torch.isfinite(r[k]).all(), k
    assert r["pred_action"].shape == (2, 21)
    g = dict(model.named_parameters())
    assert float(step.flat.grad.abs().max()) > 0
    assert g["upsample1.weight"].shape[0] == 336
