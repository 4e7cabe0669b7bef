#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel aggregated per CUDA source line (needs -lineinfo and
`ncu --import-source on`).   usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
inst, samp, text = collections.Counter(), collections.Counter(), {}
cur_file, hdr = "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_ex, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]))
    if r[2] != "-":  # a SASS row under the line: the line row already carries the totals
        continue
    inst[key] += int(r[i_ex]) if r[i_ex].isdigit() else 0
    samp[key] += int(r[i_s]) if r[i_s].isdigit() else 0
    text[key] = r[1].strip()
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
for key, n in inst.most_common(top):
    print(f"{100 * n / tot_i:5.1f}% inst {100 * samp[key] / max(tot_s, 1):5.1f}% stall  {key[0]}:{key[1]:<4d} {text[key][:110]}")
