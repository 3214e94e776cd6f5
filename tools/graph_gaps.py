"""Where the captured step's time goes between kernels: replay the CUDA graph under torch.profiler (CUPTI kernel records),
then report the sum of kernel durations, the span from the first kernel's start to the last one's end, and the idle gaps
by the kernel that precedes them.     python tools/graph_gaps.py  -> gpurun_out/graph_gaps.json (run on the GPU box)"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200")]


def main():
    import bench
    from b200caps.step import StepArgs, TrainStep
    from models.capsules_ucf101 import CapsNet
    dev = torch.device("cuda", 0)
    torch.manual_seed(47)
    model = CapsNet(pt_path=None).to(dev)
    step = TrainStep(model, StepArgs(bv=True, gv=False, n_frames=5, wt_cons=0.1, lr=1e-4))
    hb = bench.synthetic_host_batch(8, 8, seed=47, num_classes=24)
    db = [hb[k].to(dev) for k in ("data", "fl_data", "action", "seg")]
    step.capture(16, hb["labels"], epoch=1, init_batch=db)
    step.replay(*db)
    for _ in range(5):
        step.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step.replay()
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start),
                key=lambda e: e.time_range.start)
    ks = [(e.name, e.time_range.start, e.time_range.end) for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    # the last full replay: kernels after the second-to-last adam_kernel
    adam = [i for i, k in enumerate(ks) if "adam_kernel" in k[0]]
    seg = ks[adam[-2] + 1:adam[-1] + 1]
    def short(n):
        n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        return n.split("(")[0][:56]
    by = collections.defaultdict(lambda: [0, 0.0])
    for n, s0, e0 in seg:
        by[short(n)][0] += 1
        by[short(n)][1] += e0 - s0
    busy = sum(e - s for _, s, e in seg)
    span = seg[-1][2] - seg[0][1]
    gaps = collections.defaultdict(lambda: [0, 0.0])
    overl = 0.0
    for (n0, s0, e0), (n1, s1, e1) in zip(seg, seg[1:]):
        g = s1 - e0
        key = short(n0)
        if g > 0:
            gaps[key][0] += 1
            gaps[key][1] += g
        else:
            overl += -g
    out = {"kernels": len(seg), "busy_us": busy, "span_us": span, "idle_us": sum(v[1] for v in gaps.values()), "overlap_us": overl,
           "busy_by_kernel_us": {k: {"n": v[0], "us": round(v[1], 1)} for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]},
           "gaps_by_preceding_kernel_us": {k: {"n": v[0], "us": round(v[1], 1)} for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "graph_gaps.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:6000])


if __name__ == "__main__":
    main()
