"""Pin the restatement (oracle/restate.py) against the reference's own code and write
the golden fixtures under tests/golden/.  Runs ONLY in the build container, where
/root/reference exists.   python -m oracle.make_golden [--quick]

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, restate  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


class InjectedDropout(torch.nn.Module):
    """Stands in for nn.Dropout3d so the reference consumes OUR masks, in draw order."""

    def __init__(self, masks):
        super().__init__()
        self.masks = list(masks)
        self.i = 0

    def forward(self, x):
        if not self.training:
            return x
        m = self.masks[self.i]
        self.i += 1
        return x * m.to(x.dtype)


def mask_inputs():
    g = torch.Generator().manual_seed(11)
    pm = torch.randn((2, 1, 8, 224, 224), generator=g) * 0.4
    fm = torch.randn((2, 1, 8, 224, 224), generator=g) * 0.4
    return pm, fm


def sample_idx(n, k, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (k,), generator=g)


def summarize(t: torch.Tensor, k=64, seed=1):
    f = t.detach().double().reshape(-1)
    idx = sample_idx(f.numel(), min(k, f.numel()), seed)
    return dict(shape=list(t.shape), sum=float(f.sum()), abssum=float(f.abs().sum()), maxabs=float(f.abs().max()),
                idx=idx.tolist(), vals=f[idx].tolist())


def build_ref_model(ns, sd64, masks=None):
    model = ns.caps.CapsNet()
    model = model.double()
    model.conv_caps.ln_2pi = model.conv_caps.ln_2pi.double()
    missing = model.load_state_dict(sd64, strict=True)
    if masks is not None:
        model.dropout3d = InjectedDropout(masks)
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--big", type=int, default=0, help="only write the N+N-clip full-step goldens (step_NpN.json)")
    args = ap.parse_args()
    if args.big:
        return step_big(args.big)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ns = ref_import.import_reference(double=True)
    report = {}

    # ---- 1. state_dict contract --------------------------------------------------
    ref_model = ns.caps.CapsNet()
    ref_sd = ref_model.state_dict()
    spec = restate.state_dict_spec(24)
    assert [k for k, _, _ in spec] == list(ref_sd.keys()), "state_dict key order differs"
    for k, shp, _ in spec:
        assert tuple(ref_sd[k].shape) == tuple(shp), (k, ref_sd[k].shape, shp)
    with open(os.path.join(GOLD, "state_dict_spec.json"), "w") as f:
        json.dump([[k, list(s)] for k, s, _ in spec], f)
    report["state_dict_keys"] = len(spec)
    report["n_params"] = int(sum(p.numel() for p in ref_model.parameters()))
    del ref_model

    sd64 = restate.make_state_dict(24, seed=0, dtype=torch.float64)

    # ---- 2. small-op known answers (reference functions executed directly) --------
    kat = {}
    g = torch.Generator().manual_seed(5)
    # SpreadLoss (utils/losses.py:14-37)
    x = torch.rand((5, 24), generator=g, dtype=torch.float64)
    tgt = torch.randint(0, 24, (5, 1), generator=g).float()
    l_ref, a_ref = ns.losses.SpreadLoss(num_class=24, m_min=0.2, m_max=0.9)(x, tgt)
    l_res, a_res = restate.spread_loss(x, tgt)
    assert abs(float(l_ref) - float(l_res)) < 1e-12 and abs(float(a_ref) - float(a_res)) < 1e-12
    kat["spread"] = dict(x=x.tolist(), target=tgt.view(-1).tolist(), loss=float(l_ref), absloss=float(a_ref))
    # DiceLoss (utils/losses.py:44-57) + weighted mse (:74-76)
    lg = torch.randn((2, 1, 2, 6, 6), generator=g, dtype=torch.float64)
    tt = (torch.rand((2, 1, 2, 6, 6), generator=g) > 0.6).double()
    d_ref = ns.losses.DiceLoss()(lg, tt)
    assert abs(float(d_ref) - float(restate.dice_loss(lg, tt))) < 1e-13
    kat["dice"] = dict(logits=lg.tolist(), targets=tt.tolist(), loss=float(d_ref))
    # EM routing (capsules_ucf101.py:290-309) on 6 locations
    cc = ns.caps.ConvCaps(32, 24, (1, 1), 4, stride=(1, 1), iters=3).double()
    cc.ln_2pi = cc.ln_2pi.double()
    cc.load_state_dict({"beta_u": sd64["conv_caps.beta_u"], "beta_a": sd64["conv_caps.beta_a"],
                        "weights": sd64["conv_caps.weights"]})
    xin = torch.cat([torch.randn((1, 2, 3, 512), generator=g, dtype=torch.float64) * 0.7,
                     torch.rand((1, 2, 3, 32), generator=g, dtype=torch.float64)], dim=-1).requires_grad_(True)
    out_ref = cc(xin)
    mu, a = restate.em_routing(xin[..., :512].reshape(6, 32, 16), xin[..., 512:].reshape(6, 32),
                               sd64["conv_caps.weights"][0], sd64["conv_caps.beta_u"], sd64["conv_caps.beta_a"])
    out_res = torch.cat([mu.reshape(1, 2, 3, 384), a.reshape(1, 2, 3, 24)], dim=-1)
    e = rel(out_res, out_ref)
    assert e < 1e-10, e
    report["routing_restate_vs_ref"] = e
    gw = torch.randn(out_ref.shape, generator=g, dtype=torch.float64)
    (gin_ref,) = torch.autograd.grad((out_ref * gw).sum(), xin)
    kat["routing"] = dict(x=xin.detach().reshape(6, 544).tolist(), out=out_ref.detach().reshape(6, 408).tolist(),
                          gout=gw.reshape(6, 408).tolist(), gin=gin_ref.reshape(6, 544).tolist())
    with open(os.path.join(GOLD, "kat_small.json"), "w") as f:
        json.dump(kat, f)

    # consistency masks (utils/helpers.py:8-95): reference on float32 inputs like the real call.
    # Inputs come from a dedicated generator (seed 11) so tests regenerate them bit-identically.
    pm, fm = mask_inputs()
    masks_gold = {}
    for name, fn_ref, fn_res in (
            ("bv3", lambda: ns.helpers.measure_pixelwise_var_v2(pm, fm, frames_cnt=3),
             lambda: restate.pixelwise_var_mask(pm, fm, 3)),
            ("bv5", lambda: ns.helpers.measure_pixelwise_var_v2(pm, fm, frames_cnt=5),
             lambda: restate.pixelwise_var_mask(pm, fm, 5)),
            ("bv5_sig", lambda: ns.helpers.measure_pixelwise_var_v2(pm, fm, frames_cnt=5, use_sig_output=True),
             lambda: restate.pixelwise_var_mask(pm, fm, 5, True)),
            ("gv", lambda: ns.helpers.measure_pixelwise_gradient(pm), lambda: restate.pixelwise_grad_mask(pm)),
            ("gv_thr", lambda: ns.helpers.measure_pixelwise_gradient(pm.clone(), 0.45, 0.55),
             lambda: restate.pixelwise_grad_mask(pm, 0.45, 0.55))):
        m_ref, m_res = fn_ref(), fn_res()
        assert tuple(m_ref.shape) == tuple(m_res.shape), (name, m_ref.shape, m_res.shape)
        e = float((m_ref - m_res.double()).abs().max())
        assert e < 2e-6, (name, e)
        report[f"mask_{name}_restate_vs_ref_maxabs"] = e
        masks_gold[name] = summarize(m_ref, 512, 11)
    with open(os.path.join(GOLD, "masks.json"), "w") as f:
        json.dump(masks_gold, f)

    # ---- 3. CapsNet forward, train + eval, fp64, B=2 -----------------------------
    batch = restate.synthetic_batch(1, 1, seed=47, dtype=torch.float64)
    masks = restate.make_drop_masks(2, seed=3, count=4, dtype=torch.float64)
    model = build_ref_model(ns, sd64, masks[:2])
    model.train()
    t0 = time.time()
    o_ref, a_ref, f_ref = model(batch["data"], batch["action"], batch["labels"], 1, 11)
    t_ref = time.time() - t0
    bn = restate.BNState(True)
    o_res, a_res, f_res = restate.capsnet_forward(sd64, batch["data"], batch["action"], batch["labels"], 1, 11, True,
                                                  masks[:2], bn)
    report["capsnet_train_fwd"] = dict(logits=rel(o_res, o_ref), act=rel(a_res, a_ref), feat=rel(f_res, f_ref),
                                       ref_seconds=t_ref)
    assert report["capsnet_train_fwd"]["logits"] < 1e-9 and report["capsnet_train_fwd"]["act"] < 1e-7, report
    # BN running stats restated == reference's
    new_sd = model.state_dict()
    worst = 0.0
    for p, (rm, rv) in bn.updates.items():
        worst = max(worst, rel(rm, new_sd[p + ".bn.running_mean"]), rel(rv, new_sd[p + ".bn.running_var"]))
    report["bn_running_stats"] = worst
    assert worst < 1e-10, worst
    gold_fwd = dict(train=dict(logits=summarize(o_ref), act=a_ref.tolist(), feat=summarize(f_ref),
                               bn_running={p: [summarize(rm, 8), summarize(rv, 8)] for p, (rm, rv) in
                                           list(bn.updates.items())[:4]}))
    # eval mode (argmax pose mask branch, capsules_ucf101.py:473-479); fresh model so running stats are init
    model = build_ref_model(ns, sd64, None)
    model.eval()
    with torch.no_grad():
        o_ref, a_ref, f_ref = model(batch["data"], batch["action"], batch["labels"], 0, 0)
        o_res, a_res, f_res = restate.capsnet_forward(sd64, batch["data"], batch["action"], batch["labels"], 0, 0,
                                                      False, None)
    report["capsnet_eval_fwd"] = dict(logits=rel(o_res, o_ref), act=rel(a_res, a_ref))
    assert report["capsnet_eval_fwd"]["logits"] < 1e-9, report
    gold_fwd["eval"] = dict(logits=summarize(o_ref), act=a_ref.tolist(),
                            mask_pos=int((o_ref > 0).sum()), argmax=a_ref.argmax(1).tolist())
    with open(os.path.join(GOLD, "capsnet_fwd_b2.json"), "w") as f:
        json.dump(gold_fwd, f)

    # ---- 4. the full step through the reference's train_model_interface ----------
    if not args.quick:
        _stub_main_deps()
        saved = list(sys.path)
        sys.path[:] = [ref_import.REF_ROOT] + [p for p in sys.path if "pi-consistency-activity-detection_b200" not in p]
        for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "datasets")]:
            del sys.modules[k]
        ds = types.ModuleType("datasets.ucf_dataloader")
        ds.UCF101DataLoader = object
        pk = types.ModuleType("datasets")
        pk.ucf_dataloader = ds
        sys.modules["datasets"] = pk
        sys.modules["datasets.ucf_dataloader"] = ds
        import importlib
        main_mod = importlib.import_module("main_ucf101")
        sys.path[:] = saved
        torch.randperm = lambda n, **k: torch.arange(n)   # the shuffle is an input, not part of the step
        step_gold = {}
        for cfg_name, flags in (("bv5", dict(bv=True, gv=False)), ("gv", dict(bv=False, gv=True)),
                                ("bv_gv", dict(bv=True, gv=True))):
            model = build_ref_model(ns, sd64, masks)
            model.train()
            main_mod.model = model
            main_mod.criterion_cls = main_mod.SpreadLoss(num_class=24, m_min=0.2, m_max=0.9)
            main_mod.criterion_seg_1 = torch.nn.BCEWithLogitsLoss()
            main_mod.criterion_seg_2 = main_mod.DiceLoss()
            a = types.SimpleNamespace(thresh_epoch=11, n_frames=5, predict_maps=False, lower_thresh=None,
                                      upper_thresh=None, bv_wt=0.5, gv_wt=0.5, wt_loc=1.0, wt_cls=1.0, wt_cons=0.1,
                                      **flags)
            lab = dict(data=batch["data"][:1], aug_data=batch["fl_data"][:1], action=batch["action"][:1],
                       loc_msk=batch["seg"][:1].double(), label_vid=batch["labels"][:1])
            unl = dict(data=batch["data"][1:], aug_data=batch["fl_data"][1:], action=batch["action"][1:],
                       loc_msk=batch["seg"][1:].double(), label_vid=batch["labels"][1:])
            wt_ramp = ns.ramp_ups.exp_rampup(100)(1)
            t0 = time.time()
            out = main_mod.train_model_interface(a, lab, unl, 1, wt_ramp)
            total_ref = out[4]
            model.zero_grad()
            total_ref.backward()
            t_ref = time.time() - t0
            grads_ref = {k: p.grad.clone() for k, p in model.named_parameters()}
            # restatement with autograd on the same weights
            sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
                   for k, v in sd64.items()}
            res = restate.train_step_losses(sdg, batch["data"], batch["fl_data"], batch["action"], batch["seg"],
                                            batch["labels"], epoch=1, thresh_epoch=11, n_frames=5, wt_cons=0.1,
                                            drop_masks=masks, **flags)
            names = [k for k, v in sdg.items() if v.requires_grad]
            gr = torch.autograd.grad(res["total"], [sdg[k] for k in names], allow_unused=True)
            worst_g = 0.0
            for k, g_ in zip(names, gr):
                worst_g = max(worst_g, rel(g_, grads_ref[k]))
            report[f"step_{cfg_name}"] = dict(total=abs(float(res["total"]) - float(total_ref)),
                                             loc=abs(float(res["loc"]) - float(out[5])),
                                             cls=abs(float(res["cls"]) - float(out[6])),
                                             cons=abs(float(res["cons"]) - float(out[7])),
                                             grad_worst_rel=worst_g, ref_seconds=t_ref)
            assert report[f"step_{cfg_name}"]["total"] < 1e-7 and worst_g < 1e-6, report[f"step_{cfg_name}"]
            step_gold[cfg_name] = dict(total=float(total_ref), loc=float(out[5]), cls=float(out[6]),
                                       cons=float(out[7]), act=out[1].tolist(), logits=summarize(out[0]),
                                       grads={k: summarize(v, 16, 7) for k, v in grads_ref.items()})
        with open(os.path.join(GOLD, "step_1p1.json"), "w") as f:
            json.dump(step_gold, f)

    with open(os.path.join(GOLD, "pinning_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


def step_big(n_each: int = 4):
    """Full-step goldens at n_each labeled + n_each unlabeled clips (fp64): the REFERENCE's train_model_interface +
    backward for UCF --bv (config 2) and --gv (config 3); the 21-class JHMDB step (config 4) through the restatement,
    because the reference does not ship that model.  -> tests/golden/step_{n}p{n}.json"""
    torch.set_num_threads(os.cpu_count())
    ns = ref_import.import_reference(double=True)
    _stub_main_deps()
    saved = list(sys.path)
    sys.path[:] = [ref_import.REF_ROOT] + [p for p in sys.path if "pi-consistency-activity-detection_b200" not in p]
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "datasets")]:
        del sys.modules[k]
    ds = types.ModuleType("datasets.ucf_dataloader")
    ds.UCF101DataLoader = object
    pk = types.ModuleType("datasets")
    pk.ucf_dataloader = ds
    sys.modules["datasets"] = pk
    sys.modules["datasets.ucf_dataloader"] = ds
    import importlib
    main_mod = importlib.import_module("main_ucf101")
    sys.path[:] = saved
    torch.randperm = lambda n, **k: torch.arange(n)   # the shuffle is an input, not part of the step
    n = n_each
    gold, report = {}, {}
    sd64 = restate.make_state_dict(24, seed=0, dtype=torch.float64)
    batch = restate.synthetic_batch(n, n, seed=47, dtype=torch.float64)
    masks = restate.make_drop_masks(2 * n, seed=3, count=4, dtype=torch.float64)
    for cfg_name, flags in (("ucf_bv5", dict(bv=True, gv=False)), ("ucf_gv", dict(bv=False, gv=True))):
        model = build_ref_model(ns, sd64, masks)
        model.train()
        main_mod.model = model
        main_mod.criterion_cls = main_mod.SpreadLoss(num_class=24, m_min=0.2, m_max=0.9)
        main_mod.criterion_seg_1 = torch.nn.BCEWithLogitsLoss()
        main_mod.criterion_seg_2 = main_mod.DiceLoss()
        a = types.SimpleNamespace(thresh_epoch=11, n_frames=5, predict_maps=False, lower_thresh=None, upper_thresh=None,
                                  bv_wt=0.5, gv_wt=0.5, wt_loc=1.0, wt_cls=1.0, wt_cons=0.1, **flags)
        part = lambda lo, hi: dict(data=batch["data"][lo:hi], aug_data=batch["fl_data"][lo:hi], action=batch["action"][lo:hi],
                                   loc_msk=batch["seg"][lo:hi].double(), label_vid=batch["labels"][lo:hi])
        t0 = time.time()
        out = main_mod.train_model_interface(a, part(0, n), part(n, 2 * n), 1, ns.ramp_ups.exp_rampup(100)(1))
        model.zero_grad()
        out[4].backward()
        t_ref = time.time() - t0
        grads_ref = {k: p.grad.clone() for k, p in model.named_parameters()}
        new_sd = model.state_dict()
        gold[cfg_name] = dict(source="reference main_ucf101.train_model_interface + backward (fp64)", total=float(out[4]),
                              loc=float(out[5]), cls=float(out[6]), cons=float(out[7]), act=out[1].tolist(),
                              logits=summarize(out[0]), mask_pos=int((out[0] > 0).sum()),
                              grads={k: summarize(v, 16, 7) for k, v in grads_ref.items()},
                              bn_running={k: summarize(new_sd[k], 8) for k in ("conv1.Conv3d_1a_7x7.bn.running_mean",
                                                                                "conv1.Conv3d_1a_7x7.bn.running_var",
                                                                                "conv1.Mixed_4f.b1b.bn.running_var")},
                              ref_seconds=t_ref)
        del model, grads_ref, out
        # the restatement agrees at this size too
        sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
               for k, v in sd64.items()}
        res = restate.train_step_losses(sdg, batch["data"], batch["fl_data"], batch["action"], batch["seg"], batch["labels"],
                                        epoch=1, thresh_epoch=11, n_frames=5, wt_cons=0.1, drop_masks=masks, **flags)
        report[cfg_name] = {k: abs(float(res[k]) - gold[cfg_name][k]) for k in ("total", "loc", "cls", "cons")}
        assert max(report[cfg_name].values()) < 1e-7, report
        del res, sdg
        print(cfg_name, gold[cfg_name]["total"], report[cfg_name], f"{t_ref:.0f}s", flush=True)
    # JHMDB-21 (config 4): restatement only
    sd21 = restate.make_state_dict(21, seed=0, dtype=torch.float64)
    b21 = restate.synthetic_batch(n, n, seed=47, num_classes=21, dtype=torch.float64)
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd21.items()}
    res = restate.train_step_losses(sdg, b21["data"], b21["fl_data"], b21["action"], b21["seg"], b21["labels"], epoch=1,
                                    thresh_epoch=11, n_frames=5, wt_cons=0.1, drop_masks=masks, bv=True, gv=False, num_classes=21)
    names = [k for k, v in sdg.items() if v.requires_grad]
    gr = torch.autograd.grad(res["total"], [sdg[k] for k in names], allow_unused=True)
    gold["jhmdb_bv5"] = dict(source="oracle/restate.py (the reference does not ship models/capsules_jhmdb_semi_sup_pa.py)",
                             total=float(res["total"]), loc=float(res["loc"]), cls=float(res["cls"]), cons=float(res["cons"]),
                             act=res["pred_action"].tolist(), logits=summarize(res["output"]),
                             mask_pos=int((res["output"] > 0).sum()),
                             grads={k: summarize(g_, 16, 7) for k, g_ in zip(names, gr)})
    with open(os.path.join(GOLD, f"step_{n}p{n}.json"), "w") as f:
        json.dump(gold, f)
    with open(os.path.join(GOLD, f"pinning_report_{n}p{n}.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


def _stub_main_deps():
    tv = types.ModuleType("torchvision")
    tv.datasets = types.ModuleType("torchvision.datasets")
    tv.transforms = types.ModuleType("torchvision.transforms")
    for n in ("torchvision", "torchvision.datasets", "torchvision.transforms"):
        sys.modules.setdefault(n, getattr(tv, n.split(".")[-1], tv) if "." in n else tv)


if __name__ == "__main__":
    main()
