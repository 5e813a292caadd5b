#!/usr/bin/env python3
"""`make EXTRA="-Xptxas -v" 2>&1 | python profiles/ptxas_table.py` -> one line per kernel: stack frame, spill stores, registers."""
import re
import subprocess
import sys

name = None
out = {}
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void mpcf::", "").replace("mpcf::", "")
        name = re.sub(r", double const\*.*|, mpcf::EeArgs.*|, int, double.*", ">", name)
        out[name] = [0, 0, 0]
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m and name:
        out[name][0], out[name][1] = int(m.group(1)), int(m.group(2))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        out[name][2] = int(m.group(1))
        name = None
for k in sorted(out):
    print("%-90s stack %6d  spill_st %6d  regs %3d" % (k[:90], *out[k]))
