"""Data-parallel gradient exchange of the fused step: the model shards by clip, the only
collective is the all-reduce of the flat fp32 gradient buffer, issued in buckets so the largest part overlaps the
encoder backward.  Device-agnostic host logic (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bucket_ranges(offsets: dict, numels: dict, n_total: int, first_prefix: str = "conv1.") -> List[Tuple[int, int]]:
    """Two buckets over the flat buffer (registration order = encoder first):
    [0, enc_end) = encoder parameters, [enc_end, n_total) = capsule head + decoder (ready first in backward)."""
    enc_end = 0
    for k, o in offsets.items():
        if k.startswith(first_prefix):
            enc_end = max(enc_end, o + numels[k])
    enc_end = (enc_end + 3) // 4 * 4
    return [(0, enc_end), (enc_end, n_total)]


class GradBuckets:
    def __init__(self, flat_grad: torch.Tensor, ranges: Sequence[Tuple[int, int]], group=None):
        self.flat, self.ranges, self.group = flat_grad, list(ranges), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.comm_stream: Optional[torch.cuda.Stream] = torch.cuda.Stream() if (flat_grad.is_cuda and self.world > 1) else None

    def allreduce(self, i: int):
        """SUM all-reduce of bucket i.  On CUDA it is enqueued on the side stream after everything already enqueued
        on the current stream (so the caller may keep launching backward kernels); call join() before the optimiser."""
        lo, hi = self.ranges[i]
        if self.world == 1 or hi <= lo:
            return
        if self.comm_stream is None:
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def join(self):
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
