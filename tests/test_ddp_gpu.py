"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`):
the data-parallel fused step on hardware.  SURVEY section 4 item 4: the all-reduced flat gradient / world must equal the mean
of the per-rank single-GPU gradients, and after one optimiser step every rank must hold the same weights."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from b200caps import engine, ops
        from b200caps.step import StepArgs, TrainStep
        from models.capsules_ucf101 import CapsNet
        from oracle import restate
        sd = restate.make_state_dict(24, seed=0)
        solo = [dist.new_group([r]) for r in range(world)][rank]          # 1-rank group: the single-GPU step of this rank
        batches = [restate.synthetic_batch(1, 1, seed=47 + r) for r in range(world)]
        g = torch.Generator().manual_seed(5)
        m832 = ((torch.rand((4, 832), generator=g) < 0.5).float() * 2).to(dev)
        m128 = ((torch.rand((4, 128), generator=g) < 0.5).float() * 2).to(dev)
        engine.STATE.dropout_source = lambda n, c, d: (m832 if c == 832 else m128)
        ops.set_deterministic(True)

        def grads_of(batch, group, lr):
            model = CapsNet(pt_path=None)
            model.load_state_dict(sd)
            model = model.to(dev).train()
            step = TrainStep(model, StepArgs(bv=True, n_frames=5, wt_cons=0.1, lr=lr), process_group=group)
            step(*[batch[k].to(dev) for k in ("data", "fl_data", "action", "seg")], batch["labels"], epoch=1)
            torch.cuda.synchronize()
            return step.flat.grad.clone(), step.flat.data.clone(), step.world

        # every rank computes every rank's single-GPU gradient locally
        local = []
        for r in range(world):
            gr, _, w = grads_of(batches[r], solo, 0.0)
            assert w == 1
            local.append(gr)
        mean_local = torch.stack(local).mean(0)
        # the data-parallel step: this rank's batch, NCCL all-reduce (SUM) of the flat gradient, Adam with 1 / world
        gr, weights, w = grads_of(batches[rank], None, 1e-3)
        assert w == world
        dev_rel = float((gr / world - mean_local).norm() / mean_local.norm())
        dev_max = float((gr / world - mean_local).abs().max() / mean_local.abs().max())
        # all ranks must end the step with identical weights
        ref = weights.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, weights))
        q.put((rank, dev_rel, dev_max, same))
    finally:
        engine.STATE.dropout_source = None
        ops.set_deterministic(False)
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_allreduced_gradient_is_mean_of_per_rank_gradients():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=600) for _ in ps)
    for p in ps:
        p.join(timeout=120)
    for rank, dev_rel, dev_max, same in res:
        print(f"rank {rank}: all-reduced grad / world vs mean of single-GPU grads: rel-L2 {dev_rel:.2e}, max {dev_max:.2e}; same weights {same}")
        # atomics in the weight-gradient kernels make two runs of the same step differ by fp32 rounding only
        assert dev_rel < 1e-3 and dev_max < 1e-2 and same


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_runs_under_torchrun_on_two_gpus():
    """`bench.py --gpus 2` exactly as the driver launches it (torchrun, one rank per GPU), with every leg enabled: each
    collective has to be entered by both ranks (the instrumented eager step of the kernel-timing leg all-reduces its
    gradients; run on rank 0 alone it deadlocked NCCL)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29300 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "bench.py"), "--gpus", "2", "--steps", "3", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["value"] > 0 and d["roofline"]["frac"] > 0
