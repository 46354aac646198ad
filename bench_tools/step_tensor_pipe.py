"""Time-weighted tensor-pipe activity over one step from an ncu metrics capture of one step window.
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -s <first launch of a steady-state step> -c <launches per step> --csv --log-file gpurun_out/step_metrics.csv \
      python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --sustain-s 0
  python bench_tools/step_tensor_pipe.py gpurun_out/step_metrics.csv "<command>" > profiles/r2_step_tensor_pipe.json"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
ix = {h: i for i, h in enumerate(rows[hi])}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(rows[hi]):
        continue
    d = per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
# keep ONE training step: the launches from one im2col (first kernel of a step) up to the next, the window must hold attn_bwd
seq = list(per.values())
starts = [i for i, d in enumerate(seq) if "im2col" in d["name"]]
for a, b in zip(starts, starts[1:]):
    if any("attn_bwd" in d["name"] for d in seq[a:b]):
        seq = seq[a:b]
        break
agg = collections.OrderedDict()
tot_t = tot_w = 0.0
for d in seq:
    t = d.get("gpu__time_duration.sum", 0.0) / 1e3
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("aimb::", "")
    name = re.sub(r"<.*", "", name) if not name.startswith("at::") else "at:: (torch elementwise / optimizer / head)"
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += t * tp; a[3] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot_t += t; tot_w += t * tp
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1],
       "metric": "sum_k(duration_k * sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active_k) / sum_k duration_k over the launches of one step",
       "launches": len(seq), "step_kernel_time_ms": tot_t / 1e3, "tensor_pipe_active_pct_time_weighted": tot_w / tot_t if tot_t else None,
       "per_kernel": [{"kernel": k, "launches": n, "time_us": round(t, 1), "share": round(t / tot_t, 4),
                       "tensor_pipe_active_pct": round(w / t, 1) if t else 0.0, "dram_MB_per_launch": round(b / n / 1e6, 1)}
                      for k, (n, t, w, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]) if t / tot_t > 0.001]}
print(json.dumps(out, indent=1))
