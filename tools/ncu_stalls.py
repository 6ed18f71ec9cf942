"""Summarise an `ncu --set full --import-source on` report per launch: headline metrics, stall-reason totals and the
instructions with the most stall samples (CPU only; needs the ncu CLI).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "smsp__inst_executed.sum"]
launches = rows[2:]
for li, r in enumerate(launches):
    print(f"== launch {li}: {r[idx['Kernel Name']][:110]}")
    for k in KEYS:
        if k in idx:
            print(f"   {k:75s} {r[idx[k]]:>16s} {units[idx[k]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for si, sec in enumerate(secs):
    h = sec["hdr"]
    ix = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    for r in sec["rows"]:
        for s in stalls:
            if r[ix[s]]:
                tot[s] += int(r[ix[s]])
    S = sum(tot.values()) or 1
    print(f"== source {si}: {sec['name'][:110]}\n   samples {S}: " + ", ".join(f"{k[6:]} {100 * v / S:.0f}%" for k, v in tot.most_common(8)))
    order = sorted(range(len(sec["rows"])), key=lambda i: -int(sec["rows"][i][ix["# Samples"]] or 0))[:topn]
    for i in sorted(order):
        r = sec["rows"][i]
        st = sorted(((s, int(r[ix[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[0]
        print(f"   #{i:5d} samples {r[ix['# Samples']]:>5s} exec {r[ix['Instructions Executed']]:>8s}  {r[ix['Source']].strip()[:72]:72s} {st[0][6:]}")
