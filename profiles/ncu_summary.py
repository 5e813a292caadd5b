#!/usr/bin/env python3
"""Summarises one kernel of an .ncu-rep: key raw metrics plus the stall-sample distribution over the instruction stream.
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
print(r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for i, h in enumerate(hdr):
    if h in want:
        print("%-70s %-12s %s" % (h, units[i], r[i]))
    elif "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try:
            if float(r[i].replace(",", "")) >= 0.15:
                print("%-70s %-12s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall: "), units[i], r[i]))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, row in enumerate(rows) if "Address" in row[:2])
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda row, k: int(row[ix[k]] or 0)
tot = sum(num(d, "# Samples") for d in data)
print("instructions %d, samples %d" % (len(data), tot))
keys = [k for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_no_inst", "stall_lg", "stall_mio", "stall_selected",
                    "stall_not_selected", "stall_branch_resolving", "stall_dispatch") if k in ix]
print("  ".join("%s %.1f%%" % (k[6:], 100.0 * sum(num(d, k) for d in data) / max(tot, 1)) for k in keys))
n = len(data)
for dch in range(10):
    seg = data[dch * n // 10:(dch + 1) * n // 10]
    print("decile %d: %5.1f%% of samples, long_sb %5.1f%%" % (dch, 100.0 * sum(num(d, "# Samples") for d in seg) / max(tot, 1),
                                                               100.0 * sum(num(d, "stall_long_sb") for d in seg) / max(tot, 1)))
top = sorted(enumerate(data), key=lambda t: -num(t[1], "# Samples"))[:12]
for i, d in top:
    print("%5d %6d  %s" % (i, num(d, "# Samples"), d[ix["Source"]][:100]))
