"""Data parallelism by clip: the only parallelism of the reference (mmaction/apis/train.py:106-110,
MMDistributedDataParallel -> torch DDP, 25 MB buckets).  Here it is explicit: the backbone's backward
fills one flat fp32 gradient buffer from the top (ln_post, block L-1, ...) down; every
``bucket_blocks`` blocks the finished slice is all-reduced (average) over NCCL/NVLink on a side
stream while the earlier blocks are still in backward.  One process per GPU; no other collective
exists on this path."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, group: Optional["dist.ProcessGroup"] = None, bucket_blocks: int = 3):
        self.group = group
        self.bucket_blocks = max(1, int(bucket_blocks))
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works: List = []
        self._side = None
        self.buckets_launched = 0
        self.bytes_reduced = 0

    def want_bucket(self, i: int, L: int) -> bool:
        """i = block whose gradients just completed (L = ln_post, -1 = temporal_embedding)."""
        if i == -1 or i == L:
            return False
        return (L - i) % self.bucket_blocks == 0

    def bucket_done(self, flat_grad: torch.Tensor, lo: int, hi: int):
        if self.world == 1 or hi <= lo:
            return
        sl = flat_grad[lo:hi]
        if sl.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=sl.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(sl.device))
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                w = dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            sl.record_stream(self._side)
        else:  # gloo (CPU tests): no AVG op
            w = dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._post_scale = getattr(self, "_post_scale", [])
            self._post_scale.append(sl)
        self._works.append(w)
        self.buckets_launched += 1
        self.bytes_reduced += sl.numel() * 4

    def finish(self):
        """Make the compute stream wait for every outstanding bucket."""
        for w in self._works:
            w.wait()
        self._works.clear()
        for sl in getattr(self, "_post_scale", []):
            sl.div_(self.world)
        self._post_scale = []


def allreduce_mean_(tensors: List[torch.Tensor], group=None):
    """Coalesced average of a few small tensors (the head FC grads, log scalars): one collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    o = 0
    for t in tensors:
        k = t.numel()
        t.copy_(flat[o:o + k].view_as(t))
        o += k


# ------------------------------------------------------------------------------------------------ inference sharding
# Multi-GPU test of the reference: DistributedSampler(shuffle=False) hands rank r the items r, r+W, r+2W, ... (padded by
# wrapping around so that every rank gets the same count, datasets/samplers/distributed_sampler.py:27-43), every rank
# scores its share, and collect_results_gpu (apis/test.py:159-199) pickles the per-rank lists, all_gathers the bytes,
# interleaves them back and drops the padding.  Here the payload is what it always is on this path - fp32 class scores -
# so it is one fixed-shape all_gather of [k, classes] and an interleave on the device: no pickle, no second collective.
def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of rank `rank` (rank-strided, padded by wrap-around to ceil(n_items / world) per rank)."""
    if n_items <= 0:
        return []
    per = -(-n_items // world)
    return [(rank + j * world) % n_items for j in range(per)]


def gather_scores(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """local [k, C] = scores of this rank's shard_indices(n_items, rank, world) -> [n_items, C] in dataset order, on every
    rank (apis/test.py:186-199: `ordered_results.extend(zip(*part_list))`, then `[:size]`)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local[:n_items]
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    inter = torch.stack(parts, dim=1).reshape(-1, local.shape[-1])        # item j of rank r -> position j * world + r
    return inter[:n_items]
