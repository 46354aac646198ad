#!/usr/bin/env python
"""bench.py — the AIM ViT_CLIP hot path on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]

Default workload = cfg2 of BASELINE.json, the configuration the metric is quoted on: ViT-B/16 8x224 training step, 8
clips per GPU (weak scaling; global batch 64 at N=8, the reference recipe, vitclip_base_k400.py:66-67).  A "step" = one
pass of the hot path over one batch of synthetic clips: backbone forward, I3D head, cross-entropy, adapter-only backward,
gradient all-reduce (N>1), AdamW.  cfg3 (16 frames) and cfg5 (ViT-L/14, 32 frames) are the same step on other shapes;
cfg4 is the ViT-L/14 3-view inference (videos sharded over the ranks, class scores all-gathered every step).
Clips are uint8 (as the reference's GPUNormalize pipeline delivers them, utils/module_hooks.py:35-87); the normalisation is
fused into the patch load.  Prints ONE JSON line (rank 0).

`--impl reference` times the UNMODIFIED reference backbone (vitclip_aim.py::AIM from the git-ignored baseline/_ref copy
that __graft_entry__.build() stages; the oracle port only if that copy is missing) on the host cores with all threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

NUM_CLASSES = 400
MEAN, STD = [122.769, 116.74, 104.04], [68.493, 66.63, 70.321]          # vitclip_base_k400.py:17-18
VITB = dict(input_resolution=224, patch_size=16, width=768, layers=12, heads=12)
VITL = dict(input_resolution=224, patch_size=14, width=1024, layers=24, heads=16)          # vitclip_large_k400.py:6
ADAPT = dict(drop_path_rate=0.2, num_tadapter=1, adapter_scale=0.5)                       # vitclip_base_k400.py:5-8
# TFLOP per clip (algorithmic): forward / training step
CONFIGS = {
    "cfg2": dict(model=dict(VITB, num_frames=8, **ADAPT), train=True, per_gpu=8, tflop=0.858, unit="clips/s",
                 metric="AIM ViT-B/16 8x224 train clips/s",
                 workload="AIM ViT-B/16 8x224 K400-shape training step (cfg2): 8 clips/GPU, fwd + I3D head + CE + "
                          "adapter-only bwd + grad all-reduce + AdamW"),
    "cfg3": dict(model=dict(VITB, num_frames=16, **ADAPT), train=True, per_gpu=8, tflop=1.719, unit="clips/s",
                 metric="AIM ViT-B/16 16x224 train clips/s",
                 workload="AIM ViT-B/16 16x224 training step (cfg3): 8 clips/GPU"),
    "cfg5": dict(model=dict(VITL, num_frames=32, **ADAPT), train=True, per_gpu=int(os.environ.get("AIMB200_CFG5_CLIPS", "20")),
                 tflop=15.94, unit="clips/s", metric="AIM ViT-L/14 32x224 train clips/s",
                 workload="AIM ViT-L/14 32x224 bf16 training step (cfg5), batch sized for 180 GB HBM"),
    "cfg4": dict(model=dict(VITL, num_frames=8, **ADAPT), train=False, per_gpu=8, views=3, tflop=5.60, unit="videos/s",
                 metric="AIM ViT-L/14 8x224 3-view inference videos/s",
                 workload="AIM ViT-L/14 8x224 inference (cfg4): 8 videos x 3 views per GPU per step, videos sharded over "
                          "the ranks, 'prob' average over views, class scores all-gathered every step"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return (d.get("bf16_tflops_sustained", 1389.8), d.get("bf16_tflops", 1639.4), d.get("hbm_gbs", 6542.1),
                "measured (MEASURED_PEAKS.json)")
    return 1400.0, 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index: int, period_ms: int = 20):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          str(period_ms), "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, pw, mx, reasons = [], [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                if len(f) > 6:
                    pw.append(float(f[6]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(cfg_name: str, steps: int, warmup: int, items: int):
    """One step of the same workload on the host cores, all threads.  Real reference class (vitclip_aim.py::AIM loaded
    unmodified through oracle/ref_loader.py's stand-ins for timm / mmcv) when the baseline/_ref copy or the mounted tree
    is there, else the oracle port.  Returns (items/s, s/step, cores, kind, detail)."""
    from oracle import aim_oracle as O
    from oracle import ref_loader
    C = CONFIGS[cfg_name]
    mc = C["model"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleCfg(input_resolution=mc["input_resolution"], num_frames=mc["num_frames"], patch_size=mc["patch_size"],
                      width=mc["width"], layers=mc["layers"], heads=mc["heads"])
    sd = O.fixture_state_dict(cfg)
    hw, hb = O.fixture_head(cfg, NUM_CLASSES)
    views = C.get("views", 1)
    g = torch.Generator().manual_seed(0)
    xu = torch.randint(0, 256, (items * views, 3, cfg.num_frames, 224, 224), dtype=torch.uint8, generator=g)
    mean, std = torch.tensor(MEAN).view(1, 3, 1, 1, 1), torch.tensor(STD).view(1, 3, 1, 1, 1)
    labels = torch.arange(items) % NUM_CLASSES
    if ref_loader.available():
        kind, detail = "reference", f"vitclip_aim.py::AIM ({ref_loader.source()}), torch CPU fp32"
        net = ref_loader.reference_module(cfg, sd, drop_path_rate=mc["drop_path_rate"])
        net.train(C["train"])
        head = torch.nn.Linear(cfg.width, NUM_CLASSES)
        with torch.no_grad():
            head.weight.copy_(hw), head.bias.copy_(hb)
        params = [p for p in net.parameters() if p.requires_grad] + list(head.parameters())
        fwd = lambda x: net(x)                                                   # noqa: E731
    else:
        kind, detail = "port", "oracle port (oracle/aim_oracle.py), torch CPU fp32"
        pr = {k: v.clone().requires_grad_(O.is_trainable(k) and C["train"]) for k, v in sd.items()}
        hw.requires_grad_(C["train"]), hb.requires_grad_(C["train"])
        head = lambda t: t @ hw.T + hb                                           # noqa: E731
        params = [v for v in pr.values() if v.requires_grad] + [hw, hb]
        fwd = lambda x: O.backbone(pr, x, cfg)                                   # noqa: E731
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=0.05) if C["train"] else None

    def step():
        x = (xu.float() - mean) / std                                            # GPUNormalize (module_hooks.py:35-87)
        if C["train"]:
            feat = fwd(x)
            lg = head(F.dropout(feat.mean(dim=(2, 3, 4)), 0.5, True))
            loss = F.cross_entropy(lg, labels)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return float(loss)
        with torch.no_grad():
            lg = head(fwd(x).mean(dim=(2, 3, 4)))
            return float(F.softmax(lg.view(items, views, -1), dim=2).mean(dim=1).sum())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return items * steps / dt, dt / steps, cores, kind, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C = CONFIGS[args.config]
    items = C["per_gpu"] if args.config in ("cfg2", "cfg3") else (2 if args.config == "cfg4" else 1)
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    v, spstep, cores, kind, detail = cpu_reference_rate(args.config, steps, warmup, items)
    same = items == C["per_gpu"]
    out = {"impl": "reference", "metric": C["metric"], "value": v, "unit": C["unit"], "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": C["workload"] + (" [reference arm: the same step on the host CPU]" if same else
                                                   f" [reference arm: bounded sample, {items} per step]"),
                      "per_step": items, "same_items_per_step_as_gpu_arm": same},
           "cpu_baseline": {"value": v, "unit": C["unit"], "cores": cores, "kind": kind,
                            "sample": f"{steps} steps x {items} {C['unit'].split('/')[0]} ({detail}, {cores} threads)"},
           "e2e": {"value": v, "unit": C["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
class Trainer:
    """Minimal recognizer glue around the backbone: I3D head (T-mean, dropout .5, FC), CE, AdamW, grad sync."""

    def __init__(self, device, world, model_cfg, dtype="bf16", seed=0, train=True, views=1):
        import aimb200
        self.aimb = aimb200
        torch.manual_seed(seed)
        m = aimb200.build_backbone(dict(type="ViT_CLIP", block="aim", compute_dtype=dtype, **model_cfg))
        m.init_weights()                                   # random init (no CLIP weights offline) + freeze rule
        m.set_input_normalization(MEAN, STD)               # uint8 clips in: GPUNormalize fused into the patch load
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():                              # init_weights() zeroes D_fc2 / biases / temporal_embedding:
            for n, p in m.named_parameters():              # randomise them so the adapter paths do real work
                if "D_fc2" in n or n.endswith("bias") or "temporal_embedding" in n:
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))
        self.backbone = m.to(device).train(train)
        W = model_cfg["width"]
        self.hw = (0.01 * torch.randn(NUM_CLASSES, W, generator=g)).to(device).requires_grad_(train)
        self.hb = torch.zeros(NUM_CLASSES).to(device).requires_grad_(train)   # I3DHead init (i3d_head.py:49-51)
        self.world, self.views, self.train = world, views, train
        self.sync = None
        if train:
            self.sync = aimb200.GradSync(bucket_blocks=int(os.environ.get("AIMB200_BUCKET_BLOCKS", "6"))) if world > 1 else None
            if self.sync is not None:
                self.backbone.attach_grad_sync(self.sync)
            decay = [p for n, p in self.backbone.named_parameters() if p.requires_grad and "Adapter" in n and n.endswith("weight")]
            nodecay = [p for n, p in self.backbone.named_parameters() if p.requires_grad and not ("Adapter" in n and n.endswith("weight"))]
            if os.environ.get("AIMB200_FLAT_ADAMW", "1") == "1":
                # the backbone's trainable tensors are views of one flat buffer: one AdamW kernel for all of them
                # (aimb200.FlatAdamW, same update as torch.optim.AdamW); the two head tensors stay on torch's optimizer
                head_opt = torch.optim.AdamW([{"params": [self.hw], "weight_decay": 0.05}, {"params": [self.hb], "weight_decay": 0.0}],
                                             lr=3e-4, fused=True, capturable=True)
                self.opt = aimb200.FlatAdamW(self.backbone, lr=3e-4, weight_decay=0.05, extra=head_opt)
            else:
                self.opt = torch.optim.AdamW([{"params": decay + [self.hw], "weight_decay": 0.05},
                                              {"params": nodecay + [self.hb], "weight_decay": 0.0}], lr=3e-4, fused=True,
                                             capturable=True)

    def step(self, x, labels):
        if not self.train:
            return self.infer(x, labels)
        feat = self.backbone(x)
        pooled = F.dropout(feat.mean(dim=(2, 3, 4)), 0.5, True)
        logits = F.linear(pooled, self.hw, self.hb)
        loss = F.cross_entropy(logits, labels)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.sync is not None:
            self.aimb.parallel.allreduce_mean_([self.hw.grad, self.hb.grad])
        self.opt.step()
        return loss

    @torch.no_grad()
    def infer(self, x, labels):
        """cfg4: x [videos, views, 3, T, H, W] (this rank's share) -> 'prob' average over the views
        (recognizers/base.py:186-192), scores of all ranks gathered (apis/test.py:159-199); returns the top-1 hit rate."""
        n = x.shape[0]
        feat = self.backbone(x.reshape((-1,) + x.shape[2:]))
        lg = F.linear(feat.mean(dim=(2, 3, 4)), self.hw, self.hb)
        prob = F.softmax(lg.view(n, self.views, -1), dim=2).mean(dim=1)
        allp = self.aimb.gather_scores(prob, n * self.world)
        return (allp[:n].argmax(1) == labels).float().mean()


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import aimb200
    from aimb200 import lib
    lib.load()
    C = dict(CONFIGS[args.config])
    if args.per_gpu > 0:
        C["per_gpu"] = args.per_gpu
    mc = dict(C["model"], checkpoint=bool(args.checkpoint))
    B, views, T = C["per_gpu"], C.get("views", 1), mc["num_frames"]
    tr = Trainer(dev, world, mc, dtype=args.dtype, train=C["train"], views=views)
    g = torch.Generator().manual_seed(2 + rank)
    shape = (B, views, 3, T, 224, 224) if not C["train"] else (B, 3, T, 224, 224)
    host_x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g).pin_memory()       # synthetic uint8 clips
    host_y = torch.randint(0, NUM_CLASSES, (B,), generator=g).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Device time of `steps` calls (CUDA events on the current stream), max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    use_graph = not args.no_graph
    step_fn = tr.step
    captured_launches = None
    if use_graph:
        l_before = lib.launches
        try:     # the whole step (incl. the bucketed NCCL all-reduces / the score all-gather at N>1) replayed as one CUDA graph
            graphed = aimb200.GraphedStep(tr.step, [dev_x, dev_y], warmup=3,
                                          before_capture=(lambda: tr.opt.zero_grad(set_to_none=True)) if C["train"] else None)
            captured_launches = (lib.launches - l_before) // 4          # 3 warm-up runs + the capture
            step_fn = graphed
        except Exception as e:   # noqa: BLE001  (capture unsupported by this NCCL/driver combination -> eager launches)
            if rank == 0:
                print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            use_graph = False
    if world > 1:                # every rank must take the same path
        flag = torch.tensor([1 if use_graph else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0 and use_graph:
            use_graph, step_fn = False, tr.step

    def step_resident():
        l2_flush.zero_()
        return step_fn(dev_x, dev_y)

    losses = []
    # ---- end-to-end: every step's uint8 clips come from pinned host memory and every step's loss is read back on the
    # host.  As a prefetching data loader + asynchronous logging would: the H2D copy of step k+1 runs on a copy stream while
    # step k computes, and the loss of step k-1 is read while step k runs (the reference syncs twice per iteration,
    # heads/base.py:90, recognizers/base.py:242).
    copy_stream = torch.cuda.Stream(device=dev)
    x_stage = [torch.empty_like(dev_x) for _ in range(2)]
    y_stage = [torch.empty_like(dev_y) for _ in range(2)]
    loss_pinned = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_consumed = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"k": 0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[slot])
            x_stage[slot].copy_(host_x, non_blocking=True)
            y_stage[slot].copy_(host_y, non_blocking=True)
            ev_copied[slot].record(copy_stream)

    def step_e2e():
        k = e2e_state["k"]
        slot = k % 2
        main = torch.cuda.current_stream()
        l2_flush.zero_()
        main.wait_event(ev_copied[slot])
        if use_graph:
            graphed.static_in[0].copy_(x_stage[slot], non_blocking=True)
            graphed.static_in[1].copy_(y_stage[slot], non_blocking=True)
            ev_consumed[slot].record(main)
            loss = graphed(graphed.static_in[0], graphed.static_in[1])
        else:
            loss = tr.step(x_stage[slot], y_stage[slot])
            ev_consumed[slot].record(main)
        loss_pinned[slot].copy_(loss.detach(), non_blocking=True)
        ev_done[slot].record(main)
        prefetch(1 - slot)                                   # H2D of the next step's clips, overlapped with this step
        if k > 0:
            ev_done[1 - slot].synchronize()
            losses.append(float(loss_pinned[1 - slot]))      # D2H read of the previous step's result
        e2e_state["k"] = k + 1

    for _ in range(max(3, args.warmup)):
        step_resident()
    # cost of the L2 flush alone (subtracted: it is measurement hygiene, not work)
    flush_ms = timed(lambda: l2_flush.zero_(), 10) / 10
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = lib.launches
    ms = timed(step_resident, args.steps)
    launches = captured_launches if use_graph else (lib.launches - l0) // args.steps
    clocks = sampler.stop() if sampler else None
    ms_step = ms / args.steps - flush_ms
    for ev in ev_consumed:
        ev.record(torch.cuda.current_stream())
    prefetch(0)
    ms_e2e = timed(step_e2e, args.steps) / args.steps - flush_ms
    losses.append(float(loss_pinned[(e2e_state["k"] - 1) % 2]))
    value = world * B / (ms_step / 1e3)
    e2e = world * B / (ms_e2e / 1e3)
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    # ---- sustained: the headline region is a fraction of a second at boost clocks; the same step for >= 5 s shows what the
    # part holds under its power cap (MEASURED_PEAKS.json: ~1.33 GHz median under seconds-long tensor load)
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / max(ms_step + flush_ms, 1e-3)) + 1)
        s2 = ClockSampler(local, period_ms=100) if rank == 0 else None
        ms_sus = timed(step_resident, n_sus) / n_sus - flush_ms
        ck = s2.stop() if s2 else None
        sustained = {"value": world * B / (ms_sus / 1e3), "unit": C["unit"], "ms_per_step": ms_sus, "steps": n_sus,
                     "seconds": n_sus * (ms_sus + flush_ms) / 1e3, "clocks": ck}

    # ---- roofline of the dominant kernel (tcgen05 GEMM).  The GEMM launches of one step are recorded (same operands, same
    # buffers) and replayed back to back as ONE CUDA graph: device time of the GEMMs alone, without per-launch event or
    # Python overhead (round-1 timed eager launches with an event pair each, which overstated kernel time by ~30 %).
    peak_sus, peak_burst, peak_hbm, peak_src = _peaks()
    calls = []          # (replay thunk, flop) of every tcgen05 GEMM launch of one step: gemm_nt and the paired gemm_dual_*
    orig = {n: getattr(lib, n) for n in ("gemm_nt", "gemm_dual_ncat", "gemm_dual_kcat")}

    def log_nt(a, w, out, impl=lib.IMPL_AUTO, **kw):
        calls.append((lambda: orig["gemm_nt"](a, w, out, impl=impl, **kw), 2.0 * a.shape[0] * w.shape[0] * a.shape[1]))
        return orig["gemm_nt"](a, w, out, impl=impl, **kw)

    def log_ncat(a, w1, w2, out1, out2, epi1, epi2):
        calls.append((lambda: orig["gemm_dual_ncat"](a, w1, w2, out1, out2, epi1, epi2),
                      2.0 * a.shape[0] * (w1.shape[0] + w2.shape[0]) * a.shape[1]))
        return orig["gemm_dual_ncat"](a, w1, w2, out1, out2, epi1, epi2)

    def log_kcat(a1, w1, a2, w2, out, **kw):
        calls.append((lambda: orig["gemm_dual_kcat"](a1, w1, a2, w2, out, **kw),
                      2.0 * a1.shape[0] * w1.shape[0] * (a1.shape[1] + a2.shape[1])))
        return orig["gemm_dual_kcat"](a1, w1, a2, w2, out, **kw)

    for n, f in (("gemm_nt", log_nt), ("gemm_dual_ncat", log_ncat), ("gemm_dual_kcat", log_kcat)):
        setattr(lib, n, f)
    tr.step(dev_x, dev_y)
    for n, f in orig.items():
        setattr(lib, n, f)
    torch.cuda.synchronize()
    gemm_fl = sum(fl for _, fl in calls)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for thunk, _ in calls:
            thunk()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg):
        for thunk, _ in calls:
            thunk()

    def gemm_replay():
        l2_flush.zero_()
        gg.replay()

    for _ in range(2):
        gemm_replay()
    gemm_ms = timed(gemm_replay, 5) / 5 - flush_ms
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0

    def prof(name, key=None):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.isfile(p):
            return None
        d = json.load(open(p))
        return d.get(key) if key else d

    tensor_pipe = prof("r2_step_tensor_pipe.json")
    step_tf = C["tflop"] * B / (ms_step / 1e3)
    roofline = {"bound": "tensor", "kernel": "gemm_tc4_kernel / gemm_dual_kernel (TMA + tcgen05/TMEM bf16 GEMM, fused epilogues; the paired kernel carries two nn.Linear per launch): every nn.Linear forward and dgrad",
                "achieved": achieved, "peak": peak_sus, "unit": "TFLOP/s", "frac": achieved / peak_sus,
                "peak_burst": peak_burst, "frac_of_burst_peak": achieved / peak_burst, "peak_source": peak_src,
                "traffic": prof("r2_gemm_traffic.json", "avg_dram_bytes_per_launch") or prof("r1_gemm_traffic.json", "avg_dram_bytes_per_launch"),
                "launches_per_step": len(calls), "avg_launch_us": gemm_ms * 1e3 / max(1, len(calls)),
                "timing": "the step's GEMM launches replayed back to back as one CUDA graph (CUDA events around the replay, L2 flushed before it)",
                "share_of_step": gemm_ms / ms_step,
                "share_basis": "GEMM-only graph replay time / full step graph time (the step graph overlaps weight-gradient and reduction branches with the GEMM chain)",
                "step_tflops_algorithmic": step_tf, "step_frac_of_sustained_peak": step_tf / peak_sus,
                "step_frac_of_burst_peak": step_tf / peak_burst, "step_frac_of_nominal_2250": step_tf / 2250.0,
                "tensor_pipe_time_weighted": tensor_pipe}

    if rank == 0:
        out = {"metric": C["metric"], "value": value, "unit": C["unit"], "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
               "config": {"workload": C["workload"], "name": args.config, "global_batch": world * B, "per_gpu": B,
                          "parallelism": f"dp{world}", "block": "aim", "activation_checkpointing": bool(args.checkpoint), "input": "uint8 clips, GPUNormalize fused into the patch load",
                          "cuda_graph": use_graph, "peak_hbm_gib": round(peak_mem, 1),
                          "l2": "flushed between steps (256 MiB memset, its time subtracted)"},
               "e2e": {"value": e2e, "unit": C["unit"], "h2d_bytes_per_step": host_x.numel() + host_y.numel() * 8,
                       "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
               "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "sustained": sustained,
               "loss_last": losses[-1] if losses else None}
        if world == 1 and not args.no_cpu_baseline:
            items = B if args.config in ("cfg2", "cfg3") else (2 if args.config == "cfg4" else 1)
            v_cpu, sp, cores, kind, detail = cpu_reference_rate(args.config, 1, 1, items)
            out["cpu_baseline"] = {"value": v_cpu, "unit": C["unit"], "cores": cores, "kind": kind,
                                   "sample": f"1 step (after 1 warm-up) x {items} {C['unit'].split('/')[0]} of the same workload ({detail})"}
        print(json.dumps(out), flush=True)
    # Teardown: a captured graph that contains NCCL kernels must be destroyed before the communicator, and a
    # stuck communicator teardown must never hang the launcher: hard-exit after a grace period.
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import threading as _th
        _th.Timer(20.0, lambda: os._exit(0)).start()
        step_fn = None
        if use_graph:
            graphed.graph.reset()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sustain-s", type=float, default=5.0, help="length of the extra sustained-clock loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-gpu", type=int, default=0, help="override the clips (videos) per GPU of the chosen config")
    ap.add_argument("--checkpoint", action="store_true", help="checkpoint=True: per-block activation recompute (vit_clip.py:318-319)")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
