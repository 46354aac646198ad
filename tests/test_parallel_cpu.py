"""CPU (gloo, world_size 2): the data-parallel gradient path — bucket boundaries over the flat gradient buffer,
the order of collectives, averaging — exercised without a GPU by driving ViT_CLIP._run_backward's callback
protocol with a fake engine."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aimb200
from aimb200.parallel import GradSync, allreduce_mean_


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = aimb200.ViT_CLIP(64, 4, 16, 256, 6, 4, 0.0)
        m.init_weights()
        params = dict(m.named_parameters())
        m._flatten_trainable(params)
        total = m._flat.numel()
        sync = GradSync(bucket_blocks=2)
        m.attach_grad_sync(sync)

        class FakeEngine:      # fills each block's gradient slice with (rank+1) and reports completion in backward order
            def backward(self, dfeat, W, WT, grads, on_done, **kw):
                L = m.layers
                for n_, g in grads.items():
                    if n_.startswith("ln_post"):
                        g.fill_(rank + 1.0)
                on_done(L)
                for i in reversed(range(L)):
                    for n_, g in grads.items():
                        if n_.startswith(f"transformer.resblocks.{i}."):
                            g.fill_((rank + 1.0) * (i + 1))
                    on_done(i)
                grads["temporal_embedding"].fill_(rank + 1.0)
                on_done(-1)

        m._engine = FakeEngine()
        from aimb200.engine import Dims
        d = Dims(B=1, T=4, n=17, D=256, heads=4, L=6, r=64, patch=16, res=64, kpad=768, num_tadapter=1, scale=0.5)
        m._step_ctx = ({}, {}, d)
        out = m._run_backward(torch.zeros(1, 256, 4))
        named = [n for n, p in m.named_parameters()]
        got = {n: g for n, g in zip(named, out) if g is not None}
        mean_rank = (1 + world) / 2.0
        ok = True
        for n_, g in got.items():
            if n_.startswith("transformer.resblocks."):
                i = int(n_.split(".")[2])
                want = mean_rank * (i + 1)
            else:
                want = mean_rank
            ok &= bool(torch.allclose(g, torch.full_like(g, want)))
        t = [torch.full((3,), float(rank)), torch.full((2, 2), float(rank) * 2)]
        allreduce_mean_(t)
        ok &= bool(torch.allclose(t[0], torch.full((3,), (world - 1) / 2.0)))
        q.put((rank, ok, sync.buckets_launched, sync.bytes_reduced, total * 4, len(got)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, buckets, nbytes, total_bytes, ngrads in res:
        assert ok, f"rank {rank}: averaged gradients wrong"
        assert buckets == 4            # 6 blocks / 2 per bucket = 3 buckets + the final (block 0-1 remainder / temb) flush
        assert nbytes == total_bytes   # every trainable element reduced exactly once
        assert ngrads == 3 * 4 * 6 + 3


def test_single_process_is_noop():
    s = GradSync(bucket_blocks=3)
    assert s.world == 1
    g = torch.ones(10)
    s.bucket_done(g, 0, 10)
    s.finish()
    assert s.buckets_launched == 0 and torch.equal(g, torch.ones(10))


# ---------------------------------------------------------------------------------------------- inference sharding
def test_shard_indices_match_distributed_sampler():
    from aimb200.parallel import shard_indices
    # datasets/samplers/distributed_sampler.py:27-43 (shuffle=False): indices += indices[:pad]; indices[rank::world]
    for n, world in [(10, 4), (8, 8), (3, 8), (17, 2), (1, 3)]:
        idx = list(range(n))
        total = -(-n // world) * world
        padded = (idx * (total // n + 1))[:total]
        for r in range(world):
            assert shard_indices(n, r, world) == padded[r::world]


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aimb200.parallel import gather_scores, shard_indices
        import torch.nn.functional as F
        C = 5
        g = torch.Generator().manual_seed(0)
        # (a) 7 videos over 2 ranks, scores already averaged per video
        full = torch.randn(7, C, generator=g)
        mine = full[shard_indices(7, rank, world)]
        got = gather_scores(mine, 7)
        ok = bool(torch.equal(got, full))
        # (b) 1 video x 3 views over 2 ranks (fewer videos than ranks): raw scores gathered, then 'prob' average
        raw = torch.randn(3, C, generator=g)
        mine = raw[shard_indices(3, rank, world)]
        gathered = gather_scores(mine, 3)
        prob = F.softmax(gathered.view(1, 3, C), dim=2).mean(dim=1)
        ok &= bool(torch.allclose(prob, F.softmax(raw.view(1, 3, C), dim=2).mean(dim=1)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_scores_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
