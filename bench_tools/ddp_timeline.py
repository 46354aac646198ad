"""Per-stream event timeline of ONE eager training step at N ranks: when each gradient bucket's all-reduce starts and
ends on the NCCL side stream, against the start / end of backward and of the optimizer step on the compute stream.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_tools/ddp_timeline.py > profiles/r2_ddp_timeline_nN.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import aimb200  # noqa: E402

C = bench.CONFIGS["cfg2"]
tr = bench.Trainer(dev, world, C["model"], train=True)
g = torch.Generator().manual_seed(2 + rank)
x = torch.randint(0, 256, (8, 3, 8, 224, 224), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 400, (8,), generator=g).to(dev)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
ev = {}


def mark(name, stream=None):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream or torch.cuda.current_stream())
    ev[name] = e


buckets = []
if tr.sync is not None:
    orig = tr.sync.bucket_done

    def bucket_done(flat, lo, hi):
        i = len(buckets)
        mark(f"bucket{i}_ready")                                  # compute stream: this bucket's gradients are complete
        orig(flat, lo, hi)
        mark(f"bucket{i}_allreduce_end", tr.sync._side)
        buckets.append((lo, hi))

    tr.sync.bucket_done = bucket_done
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
mark("step_start")
feat = tr.backbone(x)
logits = F.linear(F.dropout(feat.mean(dim=(2, 3, 4)), 0.5, True), tr.hw, tr.hb)
loss = F.cross_entropy(logits, y)
tr.opt.zero_grad(set_to_none=True)
mark("backward_start")
loss.backward()
mark("backward_end_incl_allreduce_wait")
if tr.sync is not None:
    aimb200.parallel.allreduce_mean_([tr.hw.grad, tr.hb.grad])
tr.opt.step()
mark("step_end")
torch.cuda.synchronize()
t0 = ev["step_start"]
out = {"world": world, "rank": rank, "note": "eager launches (no CUDA graph): absolute times are longer than the captured step, the ORDER and overlap are what this shows",
       "events_ms": {k: round(t0.elapsed_time(e), 3) for k, e in sorted(ev.items(), key=lambda kv: t0.elapsed_time(kv[1]))},
       "buckets_bytes": [(hi - lo) * 4 for lo, hi in buckets]}
if rank == 0:
    print(json.dumps(out, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
