"""Per-kernel census of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, HMMA = legacy mma.sync.
    python bench_tools/sass_census.py > profiles/r2_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "adapt-image-models_b200", "libaimb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU.EX2", "LDGSTS"]
cnt = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("aimb::", "")
        cnt[cur] = collections.Counter()
        continue
    if cur:
        for k in keys:
            if re.search(r"\b" + re.escape(k) + r"\b", line):
                cnt[cur][k] += 1
print(f"# SASS census of `{os.path.basename(so)}` ({os.path.getsize(so) / 1e6:.1f} MB, {len(cnt)} kernels): instruction counts per kernel\n")
print("| kernel | " + " | ".join(keys) + " |\n|---|" + "---:|" * len(keys))
agg = collections.OrderedDict()
for k, c in cnt.items():
    base = re.sub(r"<.*", "", k)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
for base, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["UTCHMMA"]):
    print(f"| `{base}` ({n} instantiation{'s' if n > 1 else ''}) | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
