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
