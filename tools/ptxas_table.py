#!/usr/bin/env python
"""Register / spill / stack table of every kernel of the library (ptxas -v, sm_100a; no GPU needed):

    python tools/ptxas_table.py > profiles/r01_ptxas_resources.csv
"""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mmmp_b200 import _build as B  # noqa: E402

ENTRY = re.compile(
    r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n"
    r"\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
    r"ptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")
print("# ptxas -v resource usage of every kernel of libmmmp_b200.so (sm_100a); tools/ptxas_table.py")
print("kernel,registers,spill_store_B,spill_load_B,stack_B,smem_static_B")
for src in B.SOURCES:
    r = subprocess.run([B.NVCC, *B.FLAGS, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", os.devnull],
                       capture_output=True, text=True)
    for m in ENTRY.finditer(r.stderr):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").split("(")[0]
        print(f"\"{src}: {name}\",{m.group(5)},{m.group(3)},{m.group(4)},{m.group(2)},{m.group(6) or 0}")
