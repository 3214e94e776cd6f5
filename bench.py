#!/usr/bin/env python
"""bench.py -- train clips/s of the semi-supervised step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bv|gv|bvgv]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

ours      : the fused step of b200caps (pi-consistency-activity-detection_b200/b200caps/step.py) on config 2 of
            BASELINE.json (UCF101-24, 8 labeled + 8 unlabeled clips per GPU, --bv, wt_cons 0.1, synthetic data,
            random-init weights).  Prints ONE JSON line (rank 0).
reference : the reference's own algorithm on the box's host cores (the oracle port of its training step; the
            reference is CPU-runnable Python, SURVEY config 1) on a bounded sample (1 + 1 clips per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pi-consistency-activity-detection_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "train clips/sec (labeled+unlabeled fwd+bwd)"
UNIT = "clips/s"
WORKLOAD = "UCF101-24 semi-supervised step bs 8 labeled + 8 unlabeled, --bv temporal-variance mask, wt_cons 0.1, l2, 1xB200"
# algorithmic conv / transposed-conv FLOPs per clip (both passes, fwd + dgrad + wgrad)
FLOP_PER_CLIP = 713.8e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bv", choices=["bv", "gv", "bvgv", "none"])
    ap.add_argument("--clips", type=int, default=8, help="labeled (= unlabeled) clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of as one CUDA graph")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="bf16 (default) or tf32 = fp32 activations / tcgen05 kind::tf32 operands (the reference's arithmetic)")
    ap.add_argument("--classes", type=int, default=24, choices=[24, 21], help="24 = UCF101-24 (config 2/3), 21 = JHMDB-21 (config 4)")
    ap.add_argument("--u8", action="store_true", help="uint8 input pipeline: the host hands over uint8 clips / masks, /255 and the "
                    "mirrored pass are produced on the device (26 MB instead of 180 MB per step over PCIe)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short config 3 (--gv) / config 4 (JHMDB-21) measurements")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def _run_nvml(self) -> bool:
        """Same fields through NVML (what nvidia-smi reads), sampled every 20 ms so a sub-second timed region still gets
        tens of samples; returns False when NVML is unavailable (-> nvidia-smi polling)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    idx = int(ids[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        except Exception:
            return False
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), f"{pw:.1f}"] + [("Active" if mask & b else "Not Active") for _, b in bits])
            except Exception:
                pass
            time.sleep(0.02)
        return True

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def synthetic_host_batch(n_lab, n_unl, seed, num_classes=24):
    """U[0,1) clips, flipped copies, random class, random box mask; pinned host memory."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n = n_lab + n_unl
    data = torch.rand((n, 3, 8, 224, 224), generator=g)
    action = torch.randint(0, num_classes, (n, 1), generator=g).float()
    seg = torch.zeros((n, 1, 8, 224, 224))
    for i in range(n):
        y0, x0 = [int(v) for v in torch.randint(0, 150, (2,), generator=g)]
        hh, ww = [int(v) for v in torch.randint(30, 74, (2,), generator=g)]
        seg[i, 0, :, y0:y0 + hh, x0:x0 + ww] = 1.0
    labels = torch.cat([torch.ones(n_lab), torch.zeros(n_unl)])
    fl = torch.flip(data, [4]).contiguous()
    pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
    return dict(data=pin(data), fl_data=pin(fl), action=pin(action), seg=pin(seg), labels=labels)


# ------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(steps: int, warmup: int):
    """The reference's training step (oracle port, fp32, torch CPU autograd) on 1 labeled + 1 unlabeled clip,
    all host threads.  Returns (median seconds per step, threads, min seconds, all times)."""
    import torch
    from oracle import restate
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = restate.make_state_dict(24, seed=0)
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    b = restate.synthetic_batch(1, 1, seed=47)
    masks = restate.make_drop_masks(2, seed=3, count=4)
    names = [k for k, v in sdg.items() if v.requires_grad]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = restate.train_step_losses(sdg, b["data"], b["fl_data"], b["action"], b["seg"], b["labels"], epoch=1,
                                        bv=True, gv=False, n_frames=5, wt_cons=0.1, drop_masks=masks)
        torch.autograd.grad(res["total"], [sdg[k] for k in names], allow_unused=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ts = sorted(times)
    return ts[len(ts) // 2], threads, ts[0], times


def cpu_config1_time(steps: int = 5, warmup: int = 1):
    """CapsNet().train(), batch 2, supervised BCE + Dice + Spread loss, forward + backward
    (oracle port of the reference classes), all host threads: (median s, min s, threads)."""
    import torch
    from oracle import restate
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = restate.make_state_dict(24, seed=0)
    sdg = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}
    g = torch.Generator().manual_seed(0)
    data = torch.rand((2, 3, 8, 224, 224), generator=g)
    action = torch.randint(0, 24, (2, 1), generator=g).float()
    seg = (torch.rand((2, 1, 8, 224, 224), generator=g) > 0.7).float()
    labels = torch.ones(2)
    names = [k for k, v in sdg.items() if v.requires_grad]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out, act, _ = restate.capsnet_forward(sdg, data, action, labels, 1, 11, True, None, restate.BNState(True))
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, seg) + restate.dice_loss(out, seg) + \
            restate.spread_loss(act, action, 0.2, 0.9)[0]
        torch.autograd.grad(loss, [sdg[k] for k in names], allow_unused=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ts = sorted(times)
    return ts[len(ts) // 2], ts[0], threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    sec, threads, sec_min, _ = cpu_reference_step_time(max(1, args.steps), max(0, args.warmup))
    value = 2.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic (U[0,1) clips, random-init weights)",
        "config": {"workload": WORKLOAD, "sample": "1 labeled + 1 unlabeled clip per step (bounded sample of the 8+8 workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "median_s_per_step": sec, "min_s_per_step": sec_min,
                         "sample": "oracle port of train_model_interface + backward, 1+1 clips, fp32, torch CPU, median over the timed steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a GPU: the b200caps hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from b200caps import _abi, ops, plans
    from b200caps.step import StepArgs, TrainStep
    if args.classes == 21:
        from models.capsules_jhmdb_semi_sup_pa import CapsNet
    else:
        from models.capsules_ucf101 import CapsNet
    plans.set_precision(args.precision)

    torch.manual_seed(47 + rank)
    model = CapsNet(pt_path=None).to(dev)
    sa = StepArgs(bv=args.mode in ("bv", "bvgv"), gv=args.mode in ("gv", "bvgv"), n_frames=5, wt_cons=0.1, lr=1e-4)
    step = TrainStep(model, sa)
    P = 2 * args.clips
    hb = synthetic_host_batch(args.clips, args.clips, seed=47 + rank, num_classes=args.classes)
    if args.u8:
        # what the dataloader decodes (ucf_dataloader.py:162-185): uint8 frames and masks; aug_data never exists on the host
        pin = lambda t: t.pin_memory()
        hb["data"] = pin((hb["data"] * 255.0).round().clamp_(0, 255).to(torch.uint8))
        hb["seg"] = pin(hb["seg"].to(torch.uint8))
        hb["fl_data"] = None
    db = {k: (v.to(dev) if (k != "labels" and v is not None) else v) for k, v in hb.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the step as one CUDA graph (removes ~600 launches + the Python tape from the critical path) -------
    use_graph = not args.no_graph
    if use_graph:
        step.capture(P, hb["labels"], epoch=1, init_batch=(db["data"], db["fl_data"], db["action"], db["seg"]), uint8_inputs=args.u8)
        launches_per_step = step.launches_per_step

        def run_resident():
            return step.replay()

        def run_e2e_loop(k):
            # every step's inputs come from pinned host memory inside the loop (k H2D copies for k steps) and every
            # step's loss goes back to the host; the copy of step i+1 overlaps the compute of step i (TrainStep.prefetch)
            step.prefetch(hb["data"], hb["fl_data"], hb["action"], hb["seg"])
            for i in range(k):
                r = step.replay()
                if i + 1 < k:
                    step.prefetch(hb["data"], hb["fl_data"], hb["action"], hb["seg"])
                out_host.copy_(r["total"].detach().reshape(1), non_blocking=True)
        step.replay(db["data"], db["fl_data"], db["action"], db["seg"])     # device-resident inputs for `value`
    else:
        launches_per_step = None

        def run_resident():
            return step(db["data"], db["fl_data"], db["action"], db["seg"], db["labels"], epoch=1)

        def run_e2e_loop(k):
            for _ in range(k):
                d = {kk: (hb[kk].to(dev, non_blocking=True) if hb[kk] is not None else None) for kk in ("data", "fl_data", "action", "seg")}
                r = step(d["data"], d["fl_data"], d["action"], d["seg"], hb["labels"], epoch=1)
                out_host.copy_(r["total"].detach().reshape(1), non_blocking=True)
    out_host = torch.empty(1, dtype=torch.float32).pin_memory()

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    for _ in range(args.warmup):
        run_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = run_resident()
    e1.record()
    barrier()
    launches = (_abi.launch_count() - l0) if not use_graph else launches_per_step * args.steps
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = P * world / (ms / 1e3)
    loss_val = float(res["total"])

    # ---- end-to-end through the public call with HOST buffers ------------------------------------------
    h2d = sum(hb[k].numel() * hb[k].element_size() for k in ("data", "fl_data", "action", "seg") if hb[k] is not None)
    run_e2e_loop(max(1, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = P * world / (ms_e2e / 1e3)

    # ---- kernel-level timing: the dominant family (tcgen05 implicit GEMM) and every bandwidth-class kernel ------------
    roofline = None
    if rank != 0 and not args.no_kernel_timing:
        # the instrumented eager step below all-reduces its gradients: every rank has to run it (rank 0 alone deadlocks NCCL)
        step(db["data"], db["fl_data"], db["action"], db["seg"], db["labels"], epoch=1)
        torch.cuda.synchronize()
    if rank == 0 and not args.no_kernel_timing:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32 = args.precision == "tf32"
        # tf32 tensor peak = half the bf16 one (B200_PROFILING.md: 1.1 vs 2.25 PFLOP/s nominal)
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) * (0.5 if tf32 else 1.0)
        burst = (float(peaks["bf16_tflops"]) * (0.5 if tf32 else 1.0)) if peaks.get("bf16_tflops") else None
        hbm = float(peaks.get("hbm_gbs", 6553.3))
        ops.TIMING, ops.BW_TIMING = [], []
        step(db["data"], db["fl_data"], db["action"], db["seg"], db["labels"], epoch=1)   # eager, instrumented
        torch.cuda.synchronize()
        recs, ops.TIMING = ops.TIMING, None
        bw, ops.BW_TIMING = ops.BW_TIMING, None
        t_ms = sum(a.elapsed_time(b) for _, a, b, _, _ in recs)
        flop_alg = FLOP_PER_CLIP * P * (11.40 / 11.42 if args.classes == 21 else 1.0)
        flop_exec = 2.0 * sum(m for _, _, _, m, _ in recs)
        bytes_alg = sum(nb for _, _, _, _, nb in recs)
        achieved = flop_alg / (t_ms / 1e3) / 1e12
        # DRAM traffic of the same launches from the committed ncu pass (profiles/r0N_traffic.json, written by
        # tools/summarize_profiles.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`): per launch, like
        # `achieved` (total over the step's conv launches / number of launches)
        traffic = None
        for tag in ("r02", "r01"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json")))
                if int(tj.get("clips_per_gpu", P)) == P and args.precision == tj.get("precision", "bf16") and args.classes == 24:
                    traffic = float(tj["igemm_dram_bytes_per_step"]) / max(1, int(tj.get("launches", len(recs))))
                    break
            except Exception:
                pass
        top = max(recs, key=lambda r: r[1].elapsed_time(r[2])) if recs else None
        fam = {}
        for name, a, b, nbytes in bw:
            f = fam.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
            f["launches"] += 1
            f["ms"] += a.elapsed_time(b)
            f["bytes"] += nbytes
        for f in fam.values():
            f["GBps"] = f["bytes"] / max(f["ms"], 1e-9) / 1e6
            f["frac_of_hbm"] = f["GBps"] / hbm
        bw_ms = sum(f["ms"] for f in fam.values())
        bw_bytes = sum(f["bytes"] for f in fam.values())
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, mean over the step's conv launches)",
                    "algorithmic_flop_per_launch": flop_alg / max(1, len(recs)),
                    # compulsory bytes of the same launches as they are executed (operands and results once each) next to the
                    # ncu DRAM traffic above: their ratio is the re-read overhead
                    "algorithmic_bytes_per_launch": bytes_alg / max(1, len(recs)),
                    "traffic_over_algorithmic": (traffic / (bytes_alg / max(1, len(recs)))) if traffic else None,
                    "longest_launch": {"layer": top[0], "ms": top[1].elapsed_time(top[2])} if top else None,
                    "kernel": "igemm_fprop_kernel + igemm_wgrad_kernel (all conv / transposed-conv launches)",
                    "kernel_ms_per_step": t_ms, "kernel_launches_per_step": len(recs),
                    "peak_burst": burst, "frac_of_burst": (achieved / burst) if burst else None,
                    "share_of_step": t_ms / ms, "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained" + (" x 0.5 (tf32)" if tf32 else ""))
                    if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                    "algorithmic_flop_per_step": flop_alg,
                    # the decoder tail runs as a collapsed per-clip transposed conv and the stem as a folded 2-D conv
                    # with zero-padded K: `achieved` counts the REFERENCE formulation's FLOPs (the contract's algorithmic
                    # figure); the FLOPs the tensor cores actually execute and their rate are stated beside it
                    "executed_flop_per_step": flop_exec, "executed_tflops": flop_exec / (t_ms / 1e3) / 1e12,
                    "executed_frac": flop_exec / (t_ms / 1e3) / 1e12 / peak,
                    "bandwidth_kernels": {"hbm_peak_GBps": hbm, "ms_per_step": bw_ms, "bytes_per_step": bw_bytes,
                                          "GBps": bw_bytes / max(bw_ms, 1e-9) / 1e6, "frac_of_hbm": bw_bytes / max(bw_ms, 1e-9) / 1e6 / hbm,
                                          "bytes": "algorithmic (every operand element once)", "by_kernel": fam}}
        if world == 1:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            agg, kinds = {}, {}
            for tag, a, b, _, _ in recs:
                d = agg.setdefault(tag, [0, 0.0])
                d[0] += 1
                d[1] += a.elapsed_time(b)
                k = kinds.setdefault(tag.split(" ")[0], [0, 0.0])
                k[0] += 1
                k[1] += a.elapsed_time(b)
            top = sorted(agg.items(), key=lambda kv: -kv[1][1])
            with open(os.path.join(ROOT, "gpurun_out", "igemm_kernel_times.json"), "w") as f:
                json.dump({"ms_per_step_total": t_ms, "by_kind": kinds, "by_layer_sorted": top, "bandwidth_kernels": fam}, f, indent=1)
    if world > 1:
        dist.barrier()

    # ---- the other single-GPU BASELINE configs, short runs (graph replay, device-resident inputs) ----------------------
    other = None
    if rank == 0 and world == 1 and not args.no_other_configs and use_graph and args.mode == "bv" and args.classes == 24:
        other = {}
        del step, model
        torch.cuda.empty_cache()
        for tag, mode, classes in (("config3_ucf24_gv", "gv", 24), ("config4_jhmdb21_bv", "bv", 21)):
            if classes == 21:
                from models.capsules_jhmdb_semi_sup_pa import CapsNet as Net
            else:
                from models.capsules_ucf101 import CapsNet as Net
            m2 = Net(pt_path=None).to(dev)
            s2 = TrainStep(m2, StepArgs(bv=mode == "bv", gv=mode == "gv", n_frames=5, wt_cons=0.1, lr=1e-4))
            hb2 = synthetic_host_batch(args.clips, args.clips, seed=47, num_classes=classes)
            d2 = [hb2[k].to(dev) for k in ("data", "fl_data", "action", "seg")]
            s2.capture(P, hb2["labels"], epoch=1, init_batch=d2)
            s2.replay(*d2)
            for _ in range(5):
                s2.replay()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(20):
                r2 = s2.replay()
            a1.record()
            torch.cuda.synchronize()
            t2 = a0.elapsed_time(a1) / 20
            other[tag] = {"ms_per_step": t2, "value": P / (t2 / 1e3), "unit": UNIT, "steps": 20, "warmup": 5, "loss_last_step": float(r2["total"])}
            del s2, m2
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, threads, sec_min, _ = cpu_reference_step_time(5, 1)
        c1_med, c1_min, _ = cpu_config1_time(5, 1)
        cpu_baseline = {"value": 2.0 / sec, "unit": UNIT, "cores": threads, "kind": "port", "median_s_per_step": sec, "min_s_per_step": sec_min,
                        "sample": "oracle port of the full --bv step + backward, 1 labeled + 1 unlabeled clip, fp32, 1 warm-up + 5 timed (median)",
                        "config1": {"value": 2.0 / c1_med, "unit": UNIT, "median_s_per_step": c1_med, "min_s_per_step": c1_min,
                                    "sample": "BASELINE.md section 4 config 1: supervised BCE + Dice + Spread, batch 2, forward + backward, "
                                              "1 warm-up + 5 timed (median)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic (U[0,1) 8x224x224 clips, random box masks, random-init weights)",
            "config": {"workload": WORKLOAD if world == 1 else WORKLOAD.replace("1xB200", f"{world}xB200 data-parallel, NCCL all-reduce"),
                       "clips_per_gpu": P, "mode": args.mode, "classes": args.classes, "n_frames": 5, "cuda_graph": use_graph,
                       "host_inputs": "uint8 clips + masks (device-side /255 and mirrored pass)" if args.u8 else "fp32 clips, mirrored clips, masks",
                       "l2": "working set per step (>20 GB of activations) >> 126 MB L2; no explicit flush",
                       "e2e_pipeline": "graph mode: the pinned-host -> device copy of step i+1 runs on a copy stream under the compute "
                                       "of step i (TrainStep.prefetch); every step's inputs are copied inside the timed region",
                       "parallelism": f"dp{world}", "loss_last_step": loss_val},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu_baseline, "other_configs": other,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
