#!/usr/bin/env python
"""Hottest source lines of the kernels in an ncu report (needs -lineinfo + --import-source on):
    python tools/hot_lines.py gpurun_out/r02_edge.ncu-rep [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr, kernel, per = None, None, defaultdict(list)
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Function Name":
        kernel = re.sub(r"\(.*", "", r[1])
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        per[kernel].append(r)
for kernel, rows in per.items():
    ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    ti, ts = sum(int(r[ci]) for r in rows) or 1, sum(int(r[cs]) for r in rows) or 1
    print(f"== {kernel}: {ti} warp instructions, {ts} samples")
    print("inst%  smpl%  lanes  line  source")
    for r in sorted(rows, key=lambda r: -int(r[cs]))[:top]:
        print(f"{100 * int(r[ci]) / ti:5.2f}  {100 * int(r[cs]) / ts:5.2f}  {int(r[ct]) / max(int(r[ci]), 1):5.1f}  {r[0]:>5}  {r[1].strip()[:120]}")
