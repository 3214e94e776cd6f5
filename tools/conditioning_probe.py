#!/usr/bin/env python
"""How well conditioned is the training step as a function of the number of clips per BatchNorm group?

Runs the ORACLE only (CPU, test infrastructure): the fp64 restatement of the reference step against the same
restatement with the CUDA path's bf16 rounding points (restate.emulate_bf16) and prints / stores the deviation of
losses, class activations, logits and per-tensor gradients.  This is the yardstick for what a bf16-mode implementation
can possibly agree to at a given configuration: if the reference's own maths moves by X under
operand rounding, asserting less than X on the GPU path is asserting noise.

    python tools/conditioning_probe.py --clips 1 2 4 --out profiles/r02_conditioning.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restate  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def run(n_lab, n_unl, mode, grads=True, rounding="bf16", classes=24):
    sd = restate.make_state_dict(classes, seed=0, dtype=torch.float64)
    b = restate.synthetic_batch(n_lab, n_unl, seed=47, dtype=torch.float64, num_classes=classes)
    masks = restate.make_drop_masks(n_lab + n_unl, seed=3, count=4, dtype=torch.float64)
    out = {}
    res = {}
    for tag in ("exact", "bf16"):
        sdg = {k: (v.clone().requires_grad_(grads) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
        ctx = (restate.emulate_tf32() if rounding == "tf32" else restate.emulate_bf16()) if tag == "bf16" else None
        if ctx:
            ctx.__enter__()
        try:
            r = restate.train_step_losses(sdg, b["data"], b["fl_data"], b["action"], b["seg"], b["labels"], epoch=1,
                                          bv=mode in ("bv", "bvgv"), gv=mode in ("gv", "bvgv"), n_frames=5, wt_cons=0.1, drop_masks=masks)
            g = None
            if grads:
                names = [k for k, v in sdg.items() if v.requires_grad]
                g = dict(zip(names, torch.autograd.grad(r["total"], [sdg[k] for k in names], allow_unused=True)))
        finally:
            if ctx:
                ctx.__exit__()
        res[tag] = (r, g)
    (re, ge), (rb, gb) = res["exact"], res["bf16"]
    out["losses_exact"] = {k: float(re[k]) for k in ("total", "loc", "cls", "cons")}
    out["losses_rel_dev"] = {k: abs(float(rb[k]) - float(re[k])) / (abs(float(re[k])) + 1e-30) for k in ("total", "loc", "cls", "cons")}
    out["logits_dev"] = rel(rb["output"], re["output"])
    out["logits_l2_dev"] = float((rb["output"] - re["output"]).norm() / re["output"].norm())
    out["act_dev"] = rel(rb["pred_action"], re["pred_action"])
    out["feat_dev"] = rel(rb["feat"], re["feat"])
    o = re["output"].detach()
    out["logit_range"] = [float(o.min()), float(o.max())]
    out["mask_flips"] = int(((rb["output"] > 0) != (re["output"] > 0)).sum())
    out["mask_pixels"] = int(o.numel())
    if grads:
        devs = {k: rel(gb[k], ge[k]) for k in ge if ge[k] is not None}
        groups = {"encoder": [v for k, v in devs.items() if k.startswith("conv1.")],
                  "primary_caps": [v for k, v in devs.items() if k.startswith("primary_caps.")],
                  "conv_caps": [v for k, v in devs.items() if k.startswith("conv_caps.")],
                  "decoder": [v for k, v in devs.items() if not k.startswith(("conv1.", "primary_caps.", "conv_caps."))]}
        out["grad_dev"] = {k: dict(median=sorted(v)[len(v) // 2], max=max(v)) for k, v in groups.items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--mode", default="bv")
    ap.add_argument("--no-grads", action="store_true")
    ap.add_argument("--round", default="bf16", choices=["bf16", "tf32"], help="rounding emulated at the CUDA path's rounding points")
    ap.add_argument("--classes", type=int, default=24, help="24 = UCF101-24, 21 = JHMDB-21")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    results = {}
    for n in a.clips:
        t0 = time.time()
        r = run(n, n, a.mode, grads=not a.no_grads, rounding=a.round, classes=a.classes)
        r["seconds"] = time.time() - t0
        results[f"{n}+{n}"] = r
        print(f"{n}+{n}", json.dumps(r), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
