#!/usr/bin/env python
"""Turn the scratch output of tools/profile_round.sh (gpurun_out/) into the tracked evidence
under profiles/:  python tools/summarise_profiles.py r01

  <R>_launch_list_summary.csv   per-kernel totals of the ncu launch list (shares, not absolutes)
  <R>_<kernel>_full.csv         selected `ncu --set full` metrics of one launch + hottest source lines
  <R>_bench_*.json, <R>_disagreement.txt, <R>_phase_probe.txt   copied as they are
Needs `ncu` on PATH (reads .ncu-rep files; no GPU required).
"""
import csv
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(DST, exist_ok=True)

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
       "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
       "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
       "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def launch_list():
    path = os.path.join(SRC, f"{R}_launches_raw.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(DST, f"{R}_launches_raw.csv"))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iu = hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(os.path.join(DST, f"{R}_launch_list_summary.csv"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 1000: bench.py --steps 1 --warmup 1 "
                f"--no-cpu --no-configs (1 M nodes per arm as in the bench line; 1 GPU; warm-up, timed step and e2e steps)\n"
                "# per-launch times are cold-cache/serialised: compare SHARES\nkernel,launches,total_ms,share\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"\"{k}\",{cnt[k]},{tot[k]:.3f},{tot[k] / total:.4f}\n")


def full(rep_name, out_name):
    rep = os.path.join(SRC, rep_name)
    if not os.path.exists(rep):
        return
    raw = list(csv.reader(ncu("-i", rep, "--page", "raw", "--csv").splitlines()))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    lines = [f"# ncu --set full --clock-control none --import-source on ({rep_name}); one column per captured launch"]
    names = [re.sub(r"\(.*", "", l[hdr.index("Kernel Name")]).replace("<unnamed>::", "") for l in launches]
    lines.append("metric,unit," + ",".join(f"\"{n}\"" for n in names))
    for i, h in enumerate(hdr):
        if h in RAW or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            lines.append(f"{h},{units[i]}," + ",".join(l[i].replace(",", "") for l in launches))
    src = list(csv.reader(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()))
    hdr2, rows, kernel = None, [], None
    per_kernel = defaultdict(list)
    for r in src:
        if r and r[0] == "Function Name":
            kernel = re.sub(r"\(.*", "", r[1]).replace("<unnamed>::", "")
        elif r and r[0] == "Line No":
            hdr2 = r
        elif hdr2 and len(r) == len(hdr2) and r[0] != "":
            per_kernel[kernel].append(r)
    for kernel, rows in per_kernel.items():
        ci, cs, ct = hdr2.index("Instructions Executed"), hdr2.index("# Samples"), hdr2.index("Thread Instructions Executed")
        ti, ts = sum(int(r[ci]) for r in rows) or 1, sum(int(r[cs]) for r in rows) or 1
        lines.append(f"# hottest source lines of {kernel} (share of warp instructions, share of samples, active lanes)")
        lines.append("inst_share,sample_share,lanes,line,source")
        for r in sorted(rows, key=lambda r: -int(r[cs]))[:25]:
            text = r[1].strip().replace('"', "'")
            lines.append(f"{int(r[ci]) / ti:.4f},{int(r[cs]) / ts:.4f},{int(r[ct]) / max(int(r[ci]), 1):.1f},{r[0]},\"{text[:110]}\"")
    open(os.path.join(DST, out_name), "w").write("\n".join(lines) + "\n")


launch_list()
if R == "r01":   # later rounds: tools/kernel_flops.py writes the per-kernel summaries
    full(f"{R}_knn_scan.ncu-rep", f"{R}_knn_scan_full.csv")
    full(f"{R}_edge.ncu-rep", f"{R}_edge_full.csv")
for name in (f"{R}_bench_1gpu.json", f"{R}_bench_reference.json", f"{R}_bench_2gpu.json", f"{R}_disagreement.txt",
             f"{R}_phase_probe.txt", f"{R}_configs_probe.txt", f"{R}_bench_4gpu.json", f"{R}_bench_8gpu.json"):
    p = os.path.join(SRC, name)
    if os.path.exists(p) and os.path.getsize(p) > 0:
        shutil.copy(p, os.path.join(DST, name))
print("\n".join(sorted(os.listdir(DST))))
