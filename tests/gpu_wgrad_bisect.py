"""GPU diagnostic: localise the upsample2 / upsample3 weight-gradient deviation of the teacher-forced
decoder test.  For each decoder layer L it prints, against the oracle run with the same bf16 rounding points:

  fwd   our layer input x_L and output                 vs the oracle's
  gy    our incoming gradient dL/d(out_L)              vs the oracle's
  A     OUR wgrad kernel fed the ORACLE's (x, dz)       vs the oracle's dW   -> the kernel on real data
  B     torch fp64 wgrad of OUR (x, dz)                 vs our dW            -> the kernel on our data
  C     our dW                                          vs the oracle's dW   -> what the model test sees

    python tests/gpu_wgrad_bisect.py            (needs a B200; ~1 minute)
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cl2ncdhw(t):
    return t.detach().float().permute(0, 4, 1, 2, 3).double().cpu()


def main():
    from b200caps import engine, ops
    from b200caps.plans import View
    from models.capsules_ucf101 import CapsNet
    from oracle import restate
    torch.set_num_threads(os.cpu_count())
    sd = restate.make_state_dict(24, seed=0)
    model = CapsNet(pt_path=None)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    b = restate.synthetic_batch(1, 1, seed=47)
    masks = restate.make_drop_masks(2, seed=3, count=4)
    it = iter(masks[:2])
    engine.STATE.dropout_source = lambda n, c, dev: next(it).reshape(n, c)
    try:
        x_cl, c56, c112, drop2 = model._encode(b["data"].cuda())
    finally:
        engine.STATE.dropout_source = None
    caps, rout = model._capsules(x_cl.detach())
    rout_leaf = rout.detach().requires_grad_(True)

    # record what every decoder layer's backward sees
    rec = {}
    orig = engine.cba_bwd
    names = {id(l): n for n, l in model._layers.items()}

    def spy(layer, bias, x, y, gy, relu, scale_nc, dx, accumulate=False):
        n = names[id(layer)]
        C = gy.C
        g = gy.t[..., gy.c_off:gy.c_off + C].detach().clone()
        yv = y.t[..., y.c_off:y.c_off + y.C].detach().clone() if y is not None else None
        rec[n] = dict(x=x.t[..., x.c_off:x.c_off + x.C].detach().clone(), gy=g, y=yv)
        return orig(layer, bias, x, y, gy, relu, scale_nc, dx, accumulate)

    engine.cba_bwd = spy
    try:
        out, act, feat = model._decode(rout_leaf, x_cl.detach(), c56.detach(), c112.detach(), drop2, b["action"].cuda(),
                                       b["labels"].cuda(), 1, 11)
        g = torch.Generator().manual_seed(9)
        w_o = torch.randn(out.shape, generator=g, dtype=torch.float64) / out.numel() ** 0.5
        (out * w_o.float().cuda()).sum().backward()
    finally:
        engine.cba_bwd = orig
    ours_dw = {n: getattr(model, n).weight.grad.detach().double().cpu() for n in ("upsample1", "upsample2", "upsample3", "upsample4",
                                                                                  "conv28", "conv56", "conv112")}

    # oracle with the same rounding points, with intermediate taps
    sd64 = {k: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    rout_in = rout.detach().double().cpu().requires_grad_(True)
    taps = {}
    with restate.emulate_bf16():
        o_em, a_em, f_em = restate.decode(sd64, rout_in, cl2ncdhw(x_cl)[:, :, 0], cl2ncdhw(c56), cl2ncdhw(c112), b["action"],
                                          b["labels"], 1, 11, True, masks[1].double(), taps=taps)
    loss = (o_em * w_o).sum()
    tap_names = ["u1", "cat28", "u2", "cat56", "u3", "cat112", "u4"]
    wnames = ["upsample1", "upsample2", "upsample3", "upsample4", "conv28", "conv56", "conv112"]
    grads = torch.autograd.grad(loss, [taps[n] for n in tap_names] + [sd64[n + ".weight"] for n in wnames])
    gtap = dict(zip(tap_names, grads[:len(tap_names)]))
    gw = dict(zip(wnames, grads[len(tap_names):]))
    print(f"logits ours vs oracle(bf16 roundings): {rel(out, o_em.detach()):.2e}")

    spec = {  # layer -> (oracle input tap, oracle output tap (post-ReLU), transposed?, stride, padding, output_padding)
        "upsample2": ("cat28", "u2", True, 2, 1, 1),
        "upsample3": ("cat56", "u3", True, 2, 1, 1),
        "upsample4": ("cat112", None, True, 2, 1, 1),
    }
    for n, (tin, tout, transposed, st, pd, op) in spec.items():
        r = rec[n]
        x_ours = cl2ncdhw(r["x"])
        x_or = taps[tin].detach()
        print(f"--- {n}: C (our dW vs oracle dW) max {rel(ours_dw[n], gw[n]):.2e}  L2 {l2(ours_dw[n], gw[n]):.2e}")
        print(f"    fwd  input {rel(x_ours, x_or):.2e} (L2 {l2(x_ours, x_or):.2e})")
        if tout is not None:
            y_or = taps[tout].detach()
            y_ours = cl2ncdhw(r["y"])
            gy_or = gtap[tout]
            gy_ours = cl2ncdhw(r["gy"])
            flips = int(((y_ours > 0) != (y_or > 0)).sum())
            print(f"    fwd  output {rel(y_ours, y_or):.2e}; ReLU mask flips {flips} of {y_or.numel()}")
            print(f"    gy   {rel(gy_ours, gy_or):.2e} (L2 {l2(gy_ours, gy_or):.2e})")
            dz_or = gy_or * (y_or > 0)
            dz_ours = gy_ours * (y_ours > 0)
        else:
            dz_or = None
            dz_ours = cl2ncdhw(r["gy"])
        # B: torch fp64 wgrad of OUR operands vs our dW
        wref = sd64[n + ".weight"].detach()
        xo = x_ours.clone().requires_grad_(False)
        wt = wref.clone().requires_grad_(True)
        yy = F.conv_transpose3d(xo, wt, None, stride=st, padding=pd, output_padding=op)
        (dw_torch_ours,) = torch.autograd.grad(yy, wt, dz_ours)
        print(f"    B    torch-fp64 wgrad(our x, our dz) vs our dW: max {rel(ours_dw[n], dw_torch_ours):.2e}  L2 {l2(ours_dw[n], dw_torch_ours):.2e}")
        if dz_or is not None:
            # A: our kernel on the oracle's operands
            layer = model._layers[n]
            xg = x_or.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).cuda()
            dzg = dz_or.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).cuda()
            dw = torch.zeros_like(getattr(model, n).weight)
            ops.conv_wgrad(layer.plan(xg.shape[1:4]), View(xg), View(dzg), dw, atomic=True)
            yy = F.conv_transpose3d(xg.float().permute(0, 4, 1, 2, 3).double().cpu(), wt, None, stride=st, padding=pd, output_padding=op)
            (dw_t,) = torch.autograd.grad(yy, wt, dzg.float().permute(0, 4, 1, 2, 3).double().cpu())
            print(f"    A    our kernel(oracle x, oracle dz) vs torch-fp64 on the same bf16 operands: max {rel(dw, dw_t):.2e};"
                  f" vs oracle dW: max {rel(dw, gw[n]):.2e}")
            # how much of C is explained by the operands alone: torch fp64 wgrad(our x, our dz) vs oracle dW
            print(f"    D    torch-fp64 wgrad(our x, our dz) vs oracle dW: max {rel(dw_torch_ours, gw[n]):.2e}")
            # and by bf16 rounding of the gradient alone: oracle operands, dz rounded to bf16
            print(f"    E    oracle dz rounded to bf16 vs exact dz -> dW change: max {rel(dw_t, gw[n]):.2e}")
    for n in ("upsample1", "conv28", "conv56", "conv112"):
        print(f"--- {n}: C max {rel(ours_dw[n], gw[n]):.2e}  L2 {l2(ours_dw[n], gw[n]):.2e}")


if __name__ == "__main__":
    main()
