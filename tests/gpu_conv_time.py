"""Timing probe (not a pytest file): CUDA-event time of fprop / dgrad / wgrad of one layer shape at the step's size.
   python tests/gpu_conv_time.py primarycaps | upsample1 | conv112"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200")]

SHAPES = {
    "primarycaps": (False, 832, 544, (1, 9, 9), (1, 1, 1), (1, 28, 28), 0, 32),
    "upsample1": (True, 384, 64, (1, 9, 9), (1, 1, 1), (1, 20, 20), 0, 32),
    "conv112": (False, 64, 64, (3, 3, 3), (1, 1, 1), (4, 112, 112), 1, 32),
}


def main(name):
    from b200caps import ops
    from b200caps.plans import ConvPlan, ConvSpec, View
    tr, Cin, Cout, k, s, dims, pad, N = SHAPES[name]
    dev = torch.device("cuda")
    p = (pad,) * 3 if isinstance(pad, int) else pad
    if k[0] == 1:
        p = (0, p[1], p[2])
    spec = ConvSpec(Cin, Cout, k, s, p, (0, 0, 0) if tr else p, (0, 0, 0), tr)
    plan = ConvPlan(spec, dims).to(dev)
    w = torch.randn(((Cin, Cout) if tr else (Cout, Cin)) + k, device=dev) * 0.02
    plan.pack(w, "fprop", ops.stream())
    plan.pack(w, "dgrad", ops.stream())
    x = torch.randn((N,) + dims + (spec.Cin_pad,), device=dev).bfloat16()
    y = torch.empty((N,) + tuple(plan.out_dims) + (spec.Cout_pad,), device=dev, dtype=torch.bfloat16)
    gy = torch.randn_like(y)
    gx = torch.empty_like(x)
    dw = torch.zeros_like(w)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / n

    tf = timeit(lambda: ops.conv_fprop(plan, "fprop", View(x), View(y)))
    td = timeit(lambda: ops.conv_fprop(plan, "dgrad", View(gy), View(gx)))
    tw = timeit(lambda: ops.conv_wgrad(plan, View(x), View(gy), dw, atomic=True))
    macs = plan.macs_fprop(N)
    print(f"{name} B2C_TAP_SKIP={os.environ.get('B2C_TAP_SKIP', '1')}: fprop {tf:.3f} ms ({2 * macs / tf / 1e9:.0f} TF/s alg)  dgrad {td:.3f} ms  "
          f"wgrad {tw:.3f} ms; taps fprop {[len(c.taps) for c in plan.fprop]} dgrad {[len(c.taps) for c in plan.dgrad]}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "primarycaps")
