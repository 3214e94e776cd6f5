"""Phase-timing probe for single conv layers at the bench shapes (N = 32 clips).  Build the library with
B2C_EXTRA_NVCC_FLAGS=-DB2C_PROF first; CTA 0 of every launch then prints per-warp-role cycle counts.
   python tests/gpu_layer_prof.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200"), os.path.join(ROOT, "tests")]

import torch
from b200caps import ops
from b200caps.plans import ConvPlan, ConvSpec, View

CASES = {
    # name: (transposed, Cin, Cout, k, stride, dims, pad, N, relu, bias, which)
    "up4_fprop": (True, 128, 128, (3, 3, 3), (2, 2, 2), (4, 112, 112), 1, 32, True, True, "fprop"),
    "up4_dgrad": (True, 128, 128, (3, 3, 3), (2, 2, 2), (4, 112, 112), 1, 32, False, False, "dgrad"),
    "pw64_128": (False, 64, 128, (1, 1, 1), (1, 1, 1), (8, 224, 224), 0, 32, False, False, "fprop"),
    "pw128_32": (False, 128, 32, (1, 1, 1), (1, 1, 1), (8, 224, 224), 0, 32, False, True, "fprop"),
    "c3_64_64": (False, 64, 64, (3, 3, 3), (1, 1, 1), (4, 112, 112), 1, 32, True, True, "fprop"),
    "pc_fprop": (False, 832, 544, (1, 9, 9), (1, 1, 1), (1, 28, 28), 0, 32, False, True, "fprop"),
    "pc_dgrad": (False, 832, 544, (1, 9, 9), (1, 1, 1), (1, 28, 28), 0, 32, False, False, "dgrad"),
    "inc_1x1": (False, 480, 192, (1, 1, 1), (1, 1, 1), (4, 28, 28), 0, 32, True, True, "fprop"),
    "up4_wgrad": (True, 128, 128, (3, 3, 3), (2, 2, 2), (4, 112, 112), 1, 32, False, False, "wgrad"),
    "c3_64_wgrad": (False, 64, 64, (3, 3, 3), (1, 1, 1), (4, 112, 112), 1, 32, False, False, "wgrad"),
    "pc_wgrad": (False, 832, 544, (1, 9, 9), (1, 1, 1), (1, 28, 28), 0, 32, False, False, "wgrad"),
    "inc3_wgrad": (False, 160, 320, (3, 3, 3), (1, 1, 1), (1, 28, 28), 1, 32, False, False, "wgrad"),
}


def run(name):
    tr, Cin, Cout, k, s, dims, pad, N, relu, use_bias, which = CASES[name]
    dev = torch.device("cuda")
    p3 = (pad,) * 3
    if tr:
        spec = ConvSpec(Cin, Cout, k, s, p3, (0, 0, 0), tuple(si - 1 for si in s), True)
        w = torch.randn((Cin, Cout) + k, device=dev) * 0.05
    else:
        spec = ConvSpec(Cin, Cout, k, s, p3, p3)
        w = torch.randn((Cout, Cin) + k, device=dev) * 0.05
    plan = ConvPlan(spec, dims).to(dev)
    st = ops.stream()
    plan.pack(w.contiguous(), "fprop", st)
    plan.pack(w.contiguous(), "dgrad", st)
    bias = torch.randn(Cout, device=dev) if use_bias else None
    if which == "wgrad":
        x = torch.randn((N,) + tuple(dims) + (spec.Cin_pad,), device=dev).bfloat16()
        dy = torch.randn((N,) + tuple(plan.out_dims) + (spec.Cout_pad,), device=dev).bfloat16()
        dw = torch.zeros_like(w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            e0.record()
            ops.conv_wgrad(plan, View(x), View(dy), dw, atomic=True, nsplit=0)
            e1.record()
            torch.cuda.synchronize()
        print(f"== {name}: {e0.elapsed_time(e1):.3f} ms", flush=True)
        return
    if which == "fprop":
        x = torch.randn((N,) + tuple(dims) + (spec.Cin_pad,), device=dev).bfloat16()
        y = torch.empty((N,) + tuple(plan.out_dims) + (spec.Cout_pad,), device=dev, dtype=torch.bfloat16)
    else:
        x = torch.randn((N,) + tuple(plan.out_dims) + (spec.Cout_pad,), device=dev).bfloat16()
        y = torch.empty((N,) + tuple(dims) + (spec.Cin_pad,), device=dev, dtype=torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        e0.record()
        ops.conv_fprop(plan, which, View(x), View(y), bias=bias if which == "fprop" else None, relu=relu)
        e1.record()
        torch.cuda.synchronize()
    print(f"== {name}: {e0.elapsed_time(e1):.3f} ms", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run(n)
