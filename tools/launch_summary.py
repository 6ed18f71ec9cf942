"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list -> markdown table (CPU only).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>_summary.md
"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot, cnt = collections.Counter(), collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |abcgpt::|<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((bool|int)\)", "", name)
    if name.startswith("at::") or "at::native" in name:
        name = "at:: (torch fill / copy helpers)"
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "us")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    tot[name] += v
    cnt[name] += 1
S = sum(tot.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / S:.1f} % |")
print(f"\n{sum(cnt.values())} launches, {S / 1e3:.2f} ms in total.")
