"""Bring-up probe for the tcgen05 implicit-GEMM kernels: every case runs in its own subprocess
under a timeout so a trapped / hung kernel cannot take the other cases down.
   python tests/gpu_igemm_probe.py            (driver)   |   python tests/gpu_igemm_probe.py CASE_INDEX (worker)
Compares against torch conv on the bf16-rounded operands in fp32 (cuDNN, TF32 off).
B2C_PROBE_PREC=tf32: the fp32-activation / tf32-operand mode (operands pre-rounded to tf32; bound 1e-4 instead of 2e-2)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200"), os.path.join(ROOT, "tests")]

CASES = [
    # name, transposed, Cin, Cout, k, stride, dims, pad, N
    ("1x1 64->64", False, 64, 64, (1, 1, 1), (1, 1, 1), (2, 8, 8), "same", 1),
    ("1x1 64->64 M=3000", False, 64, 64, (1, 1, 1), (1, 1, 1), (3, 25, 20), "same", 2),
    ("1x1 192->176", False, 192, 176, (1, 1, 1), (1, 1, 1), (2, 28, 28), "same", 2),
    ("1x1 16->24", False, 16, 24, (1, 1, 1), (1, 1, 1), (2, 9, 7), "same", 1),
    ("1x1 528->448 (2 N tiles)", False, 528, 448, (1, 1, 1), (1, 1, 1), (1, 28, 28), "same", 2),
    ("3x3x3 96->128", False, 96, 128, (3, 3, 3), (1, 1, 1), (2, 28, 28), "same", 2),
    ("3x3x3 16->32", False, 16, 32, (3, 3, 3), (1, 1, 1), (2, 14, 14), "same", 1),
    ("3x3x3 24->64 T=1", False, 24, 64, (3, 3, 3), (1, 1, 1), (1, 28, 28), "same", 2),
    ("3x3x3 s(2,1,1) 64->192", False, 64, 192, (3, 3, 3), (2, 1, 1), (4, 20, 20), "same", 1),
    ("stem 7x7x7 s2 3->64", False, 3, 64, (7, 7, 7), (2, 2, 2), (8, 32, 32), "same", 1),
    ("conv2d 9x9 832->544", False, 832, 544, (1, 9, 9), (1, 1, 1), (1, 12, 12), 0, 1),
    ("conv2d 3x3 p1 832->64", False, 832, 64, (1, 3, 3), (1, 1, 1), (1, 28, 28), (0, 1, 1), 1),
    ("convT2d 9x9 384->64", True, 384, 64, (1, 9, 9), (1, 1, 1), (1, 20, 20), 0, 1),
    ("convT3d s2 128->64", True, 128, 64, (3, 3, 3), (2, 2, 2), (1, 14, 14), 1, 2),
    ("convT3d s2 128->128", True, 128, 128, (3, 3, 3), (2, 2, 2), (2, 12, 12), 1, 1),
    ("proj 1x1 128->32 fp32out", False, 128, 32, (1, 1, 1), (1, 1, 1), (2, 16, 16), "same", 1),
    ("convT3d s2 128->64 (1,28,28) N=4", True, 128, 64, (3, 3, 3), (2, 2, 2), (1, 28, 28), 1, 4),
    ("convT3d s2 128->64 (2,56,56) N=2", True, 128, 64, (3, 3, 3), (2, 2, 2), (2, 56, 56), 1, 2),
    ("conv 3x3x3 192->64 p1 (2,56,56)", False, 192, 64, (3, 3, 3), (1, 1, 1), (2, 56, 56), 1, 2),
    # large-M cases: two 128-row tiles per scheduling unit (odd tile count -> a half-empty last unit)
    ("1x1 64->128 M=185955 (odd tiles)", False, 64, 128, (1, 1, 1), (1, 1, 1), (7, 161, 165), "same", 1),
    ("3x3x3 64->64 (8,112,112) N=2", False, 64, 64, (3, 3, 3), (1, 1, 1), (8, 112, 112), "same", 2),
    ("convT3d s2 128->64 (4,64,64) N=2", True, 128, 64, (3, 3, 3), (2, 2, 2), (4, 64, 64), 1, 2),
    ("convT2d 9x9 336->64 (JHMDB upsample1)", True, 336, 64, (1, 9, 9), (1, 1, 1), (1, 20, 20), 0, 1),
    # channel counts that are not multiples of 64: TMA path with a zero-filled channel tail (plans.TMA_TAIL)
    ("1x1 480->192 (Mixed_4b.b0)", False, 480, 192, (1, 1, 1), (1, 1, 1), (1, 28, 28), "same", 4),
    ("3x3x3 160->320 (Mixed_4f.b1b)", False, 160, 320, (3, 3, 3), (1, 1, 1), (1, 28, 28), "same", 4),
    ("3x3x3 144->288 (Mixed_4e.b1b)", False, 144, 288, (3, 3, 3), (1, 1, 1), (1, 28, 28), "same", 2),
    ("3x3x3 32->96 (Mixed_3c.b2b)", False, 32, 96, (3, 3, 3), (1, 1, 1), (2, 28, 28), "same", 2),
    ("1x1 512->24 (Mixed_4c.b2a)", False, 512, 24, (1, 1, 1), (1, 1, 1), (1, 28, 28), "same", 4),
]


def worker(idx):
    import torch
    import torch.nn.functional as F
    from b200caps import ops
    from b200caps import plans
    from b200caps.plans import ConvPlan, ConvSpec, View, same_pad
    tf32 = os.environ.get("B2C_PROBE_PREC", "bf16") == "tf32"
    if tf32:
        plans.set_precision("tf32")
    adt = plans.act_dtype()

    def rq(t):      # round to the operand precision of the mode
        if not tf32:
            return t.bfloat16().float()
        i = t.float().contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, tr, Cin, Cout, k, s, dims, pad, N = CASES[idx]
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(idx)
    x = rq(torch.randn((N, Cin) + dims, generator=g).to(dev))
    if tr:
        w = (torch.randn((Cin, Cout) + k, generator=g) / (Cin * 3) ** 0.5).to(dev)
        p = (pad,) * 3 if isinstance(pad, int) else pad
        op = tuple(si - 1 for si in s)
        spec = ConvSpec(Cin, Cout, k, s, p, (0, 0, 0), op, True)
    else:
        w = (torch.randn((Cout, Cin) + k, generator=g) / (Cin * k[0] * k[1] * k[2]) ** 0.5).to(dev)
        if pad == "same":
            pads = [same_pad(d, kk, ss) for d, kk, ss in zip(dims, k, s)]
        else:
            pp = (pad,) * 3 if isinstance(pad, int) else pad
            pads = [(v, v) for v in pp]
        spec = ConvSpec(Cin, Cout, k, s, tuple(p[0] for p in pads), tuple(p[1] for p in pads))
    wq = rq(w)
    bias = torch.randn(Cout, generator=g).to(dev)
    x.requires_grad_(True)
    wq.requires_grad_(True)
    if tr:
        y = F.conv_transpose3d(x, wq, bias, stride=s, padding=p, output_padding=op)
    else:
        xp = F.pad(x, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
        y = F.conv3d(xp, wq, bias, stride=s)
    gy = rq(torch.randn(y.shape, generator=g).to(dev))
    gx, gw = torch.autograd.grad(y, (x, wq), gy)

    plan = ConvPlan(spec, dims).to(dev)
    st = ops.stream()
    plan.pack(w.contiguous(), "fprop", st)
    plan.pack(w.contiguous(), "dgrad", st)

    def cl(t, cpad):
        t = t.detach().permute(0, 2, 3, 4, 1).contiguous()
        if t.shape[-1] < cpad:
            t = torch.cat([t, torch.zeros(t.shape[:-1] + (cpad - t.shape[-1],), device=dev)], -1)
        return t.to(adt).contiguous()

    res = {}
    xc = cl(x, spec.Cin_pad)
    out_fp32 = "fp32out" in name
    yc = torch.full((N,) + tuple(plan.out_dims) + (spec.Cout_pad,), float("nan"), device=dev,
                    dtype=torch.float32 if out_fp32 else adt)
    ops.conv_fprop(plan, "fprop", View(xc), View(yc), bias=bias)
    torch.cuda.synchronize()
    yref = y.detach().permute(0, 2, 3, 4, 1)
    res["fprop"] = float((yc.float()[..., :Cout] - yref).abs().max() / yref.abs().max())
    gyc = cl(gy, spec.Cout_pad)
    gxc = torch.full((N,) + tuple(dims) + (spec.Cin_pad,), float("nan"), device=dev, dtype=adt)
    ops.conv_fprop(plan, "dgrad", View(gyc), View(gxc))
    torch.cuda.synchronize()
    gxref = gx.permute(0, 2, 3, 4, 1)
    res["dgrad"] = float((gxc.float()[..., :Cin] - gxref).abs().max() / gxref.abs().max())
    for nsplit in (1, 0):
        dw = torch.zeros_like(w)
        ops.conv_wgrad(plan, View(xc), View(gyc), dw, atomic=True, nsplit=nsplit)
        torch.cuda.synchronize()
        res[f"wgrad_split{nsplit}"] = float((dw - gw).abs().max() / gw.abs().max())
    print("RESULT " + json.dumps(res))


def main():
    ok = True
    only = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 else range(len(CASES))
    for i in only:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True,
                               timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                res = json.loads(line[0][7:])
                bad = [k for k, v in res.items() if not (v < (1.5e-3 if (os.environ.get('B2C_PROBE_PREC') == 'tf32' and k in ('fprop', 'dgrad')) else 1e-4 if os.environ.get('B2C_PROBE_PREC') == 'tf32' else 2e-2))]
                print(f"[{i:2d}] {CASES[i][0]:32s} {'FAIL' if bad else 'ok  '} " +
                      " ".join(f"{k}={v:.2e}" for k, v in res.items()))
                ok &= not bad
            else:
                ok = False
                print(f"[{i:2d}] {CASES[i][0]:32s} CRASH rc={r.returncode}\n   stdout: {r.stdout[-600:]}\n   stderr: {r.stderr[-1200:]}")
        except subprocess.TimeoutExpired:
            ok = False
            print(f"[{i:2d}] {CASES[i][0]:32s} TIMEOUT")
        sys.stdout.flush()
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1].isdigit():
        worker(int(sys.argv[1]))
    else:
        sys.exit(main())
