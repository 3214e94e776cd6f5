"""N>1 host logic on CPU (gloo, world_size 2): flat parameter/gradient buffers, the two gradient buckets, and the
property the data-parallel step relies on -- all-reduced gradients / world == mean of the per-rank gradients."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200caps.ddp import GradBuckets, bucket_ranges
    g = torch.Generator().manual_seed(100 + rank)
    n = 1000
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    offsets = {"conv1.a": 0, "conv1.b": 300, "primary_caps.w": 600}
    numels = {"conv1.a": 300, "conv1.b": 297, "primary_caps.w": 400}
    ranges = bucket_ranges(offsets, numels, n)
    assert ranges == [(0, 600), (600, 1000)], ranges
    b = GradBuckets(flat, ranges)
    assert b.world == world and b.comm_stream is None
    b.allreduce(1)          # head + decoder bucket first (ready first in backward)
    b.allreduce(0)
    b.join()
    gathered = [torch.empty(n) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = torch.allclose(flat, sum(gathered)) and torch.allclose(flat / world, torch.stack(gathered).mean(0), atol=1e-6)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res


def test_flat_params_views_and_bucket_boundary():
    from b200caps.step import FlatParams

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = torch.nn.Linear(5, 3)          # 15 + 3 params, offsets padded to multiples of 4
            self.primary_caps = torch.nn.Linear(3, 2)

    m = Tiny()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    fp = FlatParams(m)
    assert list(m.state_dict().keys()) == list(before.keys())
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    for k, p in m.named_parameters():
        o = fp.offsets[k]
        assert o % 4 == 0
        assert p.data_ptr() == fp.data.data_ptr() + 4 * o and p.grad.data_ptr() == fp.grad.data_ptr() + 4 * o
    fp.data.add_(1.0)
    assert torch.allclose(m.conv1.weight, before["conv1.weight"] + 1)
