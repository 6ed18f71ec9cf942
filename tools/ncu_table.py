"""Markdown table of the headline metrics of every launch in an ncu report (CPU only; needs the ncu CLI).

    python tools/ncu_table.py gpurun_out/prof.ncu-rep > profiles/<name>.md
"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
# torch's own helper kernels (random fills, copies of the driver script) are not part of the product: dropped unless --all
if "--all" not in sys.argv:
    launches = [r for r in launches if "at::" not in r[idx["Kernel Name"]]]
METRICS = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
           ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
           ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
           ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
           ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
           ("launch__registers_per_thread", "registers / thread"), ("smsp__inst_executed_op_tma_st.sum", "TMA store instructions"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active")]
names = []
for r in launches:
    n = re.sub(r"\(CUtensorMap.*|\(const.*|\(float.*|\(__nv.*|\(long.*", "", r[idx["Kernel Name"]])
    n = re.sub(r"^void |unnamed>::|abcgpt::|<unnamed>::", "", n)
    names.append(n)
print("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|---|" + "---:|" * len(names))
for key, label in METRICS:
    if key not in idx:
        continue
    vals = []
    for r in launches:
        v = r[idx[key]]
        try:
            vals.append(f"{float(v.replace(',', '')):.2f}")
        except ValueError:
            vals.append(v)
    print(f"| {label} (`{key}`) | {units[idx[key]]} | " + " | ".join(vals) + " |")
