"""Summarises an ncu report: headline metrics + warp-stall samples aggregated by barrier-delimited phase and by opcode."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") or h.endswith("_per_warp_active.pct") and "issue_stalled" in h]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            v = r[hdr.index(w)]
            try:
                if float(v) == 0:
                    continue
            except ValueError:
                pass
            print(f"{w:90s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
phase = 0
ps, pi, ops, opi = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iS:
        continue
    s, n, text = int(r[iS] or 0), int(r[iI] or 0), r[iSrc].strip()
    tok = text.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    ps[phase] += s; pi[phase] += n; tot += s
    ops[op.split(".")[0]] += s; opi[op.split(".")[0]] += n
    if op.startswith("BAR"):
        phase += 1
print("total samples", tot, "instructions", sum(pi.values()))
for k in sorted(ps):
    print(f"phase {k}: samples {100 * ps[k] / tot:5.1f}%  inst {pi[k]}")
for k, v in ops.most_common(18):
    print(f"{k:10s} {100 * v / tot:5.1f}%  inst {opi[k]}")
