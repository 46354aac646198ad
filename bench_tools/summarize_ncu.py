"""Summaries of ncu outputs for profiles/.
  python bench_tools/summarize_ncu.py list gpurun_out/launches_r1.csv > profiles/r1_launch_list.md
  python bench_tools/summarize_ncu.py full gpurun_out/gemm_r1_final.ncu-rep > profiles/r1_gemm_full.md"""
import collections
import csv
import re
import subprocess
import sys


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('aimb::', '')
        v = float(r[ix['Metric Value']].replace(',', '')) / 1e3   # ns -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {len(data)} launches, {tot / 1e3:.1f} ms total (cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.0005:
            continue
        print(f"| `{k[:90]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")


WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in WANT if c in ix]
    print("# ncu --set full, selected raw metrics per launch\n")
    print("| kernel | " + " | ".join(c.replace('.avg', '').replace('pct_of_peak_sustained_', '%') for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('aimb::', '')
        print(f"| `{name}` | " + " | ".join(f"{r[ix[c]]} {units[ix[c]]}" for c in cols) + " |")


def traffic(path):
    """JSON consumed by bench.py's roofline.traffic: DRAM bytes (read + write) per GEMM launch of the capture."""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def mb(r, k):
        v = float(r[ix[k]].replace(',', ''))
        u = units[ix[k]].lower()
        return v * {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3}[u]

    launches = []
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('aimb::', '')
        launches.append({"kernel": name, "dram_read_MB": mb(r, 'dram__bytes_read.sum'), "dram_write_MB": mb(r, 'dram__bytes_write.sum'),
                         "time_us": float(r[ix['gpu__time_duration.sum']].replace(',', '')),
                         "tensor_pipe_active_pct": float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])})
    avg = sum((l["dram_read_MB"] + l["dram_write_MB"]) * 1e6 for l in launches) / max(1, len(launches))
    print(json.dumps({"source": " ".join(sys.argv[3:]) or path, "avg_dram_bytes_per_launch": avg, "launches": launches}, indent=1))


if __name__ == "__main__" and sys.argv[1] in ("list", "full", "traffic"):
    {"list": launch_list, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])


def stalls(path, top=40):
    """Per-SASS-instruction stall samples of the first kernel in the report (needs --import-source on)."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[ix['# Samples']] or 0) for r in data)
    print(f"# {rows[0][1][:100]}: {tot} samples, {len(data)} SASS instructions\n")
    agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in reasons}
    print("stall totals: " + ", ".join(f"{k[6:]} {100 * v / max(1, sum(agg.values())):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print("\n| # | samples | % | executed | top stall | SASS |\n|---:|---:|---:|---:|---|---|")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']] or 0))[:top]
    for i in sorted(order):
        r = data[i]
        s = int(r[ix['# Samples']] or 0)
        best = max(reasons, key=lambda k: int(r[ix[k]] or 0))
        print(f"| {i} | {s} | {100 * s / max(1, tot):.1f} | {r[ix['Instructions Executed']]} | {best[6:]} | `{r[ix['Source']].strip()[:70]}` |")


if __name__ == "__main__" and sys.argv[1] == "stalls":
    stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
