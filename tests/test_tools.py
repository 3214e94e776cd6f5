"""Host logic of the profiling helper (tools/profile_step.py): which launches of a step get a full ncu capture and
which --launch-skip ordinals address them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_capture_plan_addresses_the_longest_launches_of_the_last_step():
    import profile_step as ps
    step = [("k::igemm_fprop_kernel", 5), ("k::bn_stats_kernel", 1), ("k::igemm_fprop_kernel", 9), ("k::igemm_wgrad_kernel", 7),
            ("k::em_routing_bwd_kernel", 8), ("k::adam_kernel", 1)]
    rows, i = [], 0
    for rep in range(3):                      # three steps; the plan must pick the last complete one
        for name, ms in step:
            rows.append((i, name, (ms + rep) * 1e6))
            i += 1
    (lo, hi), caps, conv_skip, conv_count = ps.plan_captures(rows, top_kernels=3)
    assert (lo, hi) == (12, 18)
    assert conv_skip == 6 and conv_count == 3          # three conv launches per step, two steps before
    got = {(k, o) for k, o, _ in caps}
    # fprop: both launches of the step (ordinals 4 and 5 among all fprop launches); routing bwd: third launch; wgrad: third
    assert ("igemm_fprop_kernel", 4) in got and ("igemm_fprop_kernel", 5) in got
    assert ("em_routing_bwd_kernel", 2) in got and ("igemm_wgrad_kernel", 2) in got


def test_bench_reference_arm_other_ranks_exit_silently():
    """`bench.py --impl reference` under torchrun: rank 0 alone runs and prints, every other rank exits 0 without work."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-500:], r.stderr[-500:])


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """The product path has no CPU fallback: without a GPU `bench.py` fails loudly instead of timing something else."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_graph_gap_report_shape():
    """tools/graph_gaps.py's committed report (profiles/r02d_graph_busy.json): the captured step is busy > 99 % of its span and
    the implicit-GEMM kernels are its largest part."""
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", "r02d_graph_busy.json")))
    assert d["busy_us"] / d["span_us"] > 0.99 and d["idle_us"] < 0.01 * d["span_us"]
    conv = sum(v["us"] for k, v in d["busy_by_kernel_us"].items() if k.startswith("igemm_"))
    assert 0.5 < conv / d["busy_us"] < 0.8
