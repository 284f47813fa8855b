#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep into a small JSON + text for profiles/:
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_fill_hx_ncu_summary [kernel-name-regex]
Reads the report with `ncu -i ... --page raw/source --csv` (no GPU needed)."""
import collections
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
ksel = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + ksel, capture_output=True, text=True).stdout
import re
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
kcol = hdr.index("Kernel Name")
pick = [r for r in rows[2:] if len(sys.argv) <= 3 or re.search(sys.argv[3], r[kcol])]
vals = pick[0]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name, default=None):
    if name not in m:
        return default
    v, u = m[name]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return default
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
    return x * scale


want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
summ = {"report": rep, "kernel": m.get("Kernel Name", ("?",))[0]}
for w in want:
    summ[w] = num(w)
# pipe utilisation: on sm_100 `--set full` collects sm__inst_executed_pipe_<pipe>.avg.pct_of_peak_sustained_active
# (the smsp__inst_executed_pipe_*.sum names of older architectures are not valid here)
summ["pipe_pct_of_peak_sustained_active"] = {
    h[len("sm__inst_executed_pipe_"):-len(".avg.pct_of_peak_sustained_active")]: num(h)
    for h in m if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")}
rd, wr = summ.get("dram__bytes_read.sum") or 0, summ.get("dram__bytes_write.sum") or 0
summ["dram_bytes_total"] = rd + wr
stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v[0])
          for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
summ["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
json.dump(summ, open(out + ".json", "w"), indent=1)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page is one section per kernel: a "Kernel Name" line, a header line, then the instructions
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sect = starts[0]
for i in starts:
    if len(sys.argv) > 3 and re.search(sys.argv[3], rows[i][1]):
        sect = i
        break
end = min([i for i in starts if i > sect] + [len(rows)])
h2 = rows[sect + 1]
ix = {h: i for i, h in enumerate(h2)}
data = [r for r in rows[sect + 2:end] if len(r) >= len(h2)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
ops = collections.Counter()
for r in data:
    parts = r[ix["Source"]].split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
with open(out + ".txt", "w") as f:
    f.write(json.dumps({k: summ[k] for k in summ if k != "stall_cycles_per_issue"}, indent=1) + "\n")
    f.write("pipe utilisation (% of peak sustained, active cycles): " + json.dumps(summ["pipe_pct_of_peak_sustained_active"]) + "\n")
    f.write("stall cycles per issued instruction: " + json.dumps(summ["stall_cycles_per_issue"]) + "\n\n")
    f.write("executed warp-instructions by opcode (top 25):\n")
    for op, n in ops.most_common(25):
        f.write(f"  {op:14s} {n:14d}\n")
    f.write("\nhottest instructions by stall samples:\n")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:30]:
        n = int(r[ix["# Samples"]] or 0)
        f.write(f"  {100 * n / max(tot, 1):5.1f}%  exec={r[ix['Instructions Executed']]:>10}  {r[ix['Source']][:70]}\n")
print(json.dumps(summ, indent=1))
