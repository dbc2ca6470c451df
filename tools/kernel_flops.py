#!/usr/bin/env python
"""Executed FP32 work per unit of work of the hot kernels, from ncu captures (no GPU needed):

    python tools/kernel_flops.py r02     # reads gpurun_out/r02_<probe>.ncu-rep + gpurun_out/r02_<probe>_plain.log
                                         # writes profiles/r02_kernel_flops.json and profiles/r02_<probe>_full.csv

flop = 2 * (FFMA + 2 * FFMA2) + FADD + 2 * FADD2 + FMUL + 2 * FMUL2 thread-level instructions
(sm__sass_thread_inst_executed_op_f*_pred_on.sum) of ONE launch, divided by the library's own
work counters of the same launch (`COUNTERS` line of tools/kernel_probe.py run without ncu on
the same seeded input).  bench.py multiplies these per-unit figures with its live counters and
divides by live CUDA-event times: `roofline.achieved`.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|lts__t_bytes\.sum|launch__.*|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|sm__inst_executed\.sum|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|sm__cycles_active\.avg|sm__cycles_elapsed\.max|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"sm__inst_executed_pipe_.*|sm__pipe_.*cycles_active.*|sm__sass_thread_inst_executed_op_.*|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio)$")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
         "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        launches.append(d)
    return launches


def num(launch, name, default=0.0):
    if name not in launch:
        return default
    v, u = launch[name]
    try:
        return float(v.replace(",", "")) * SCALE.get(u, 1.0)
    except ValueError:
        return default


def fp32_flop(l):
    g = lambda n: num(l, f"sm__sass_thread_inst_executed_op_{n}_pred_on.sum")  # noqa: E731
    return 2 * (g("ffma") + 2 * g("ffma2")) + g("fadd") + 2 * g("fadd2") + g("fmul") + 2 * g("fmul2")


def counters_of(log):
    try:
        for line in open(log):
            if line.startswith("COUNTERS "):
                return json.loads(line[len("COUNTERS "):])
    except OSError:
        pass
    return {}


def kernel_name(l):
    return re.sub(r"\(.*", "", l["Kernel Name"][0]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def summary_csv(rep_name, launches, path):
    lines = [f"# ncu --set full + FP32-pipe metrics, --clock-control none ({rep_name}); one column per captured launch",
             "metric,unit," + ",".join(f"\"{kernel_name(l)}\"" for l in launches)]
    for h in launches[0]:
        if KEEP.match(h):
            lines.append(f"{h},{launches[0][h][1]}," + ",".join(l[h][0].replace(",", "") for l in launches))
    lines.append("derived:fp32_flop_thread_level,flop," + ",".join(f"{fp32_flop(l):.6g}" for l in launches))
    lines.append("derived:fp32_tflops,TFLOP/s," + ",".join(
        f"{fp32_flop(l) / max(num(l, 'gpu__time_duration.sum'), 1e-9) / 1e9:.4f}" for l in launches))
    open(path, "w").write("\n".join(lines) + "\n")


def main():
    R = sys.argv[1] if len(sys.argv) > 1 else "r02"
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    out = {}
    for probe in ("scan", "scan14", "edge", "configs", "fk", "accept"):
        rep = os.path.join(src, f"{R}_{probe}.ncu-rep")
        if not os.path.exists(rep):
            continue
        launches = raw_page(rep)
        if not launches:
            continue
        summary_csv(os.path.basename(rep), launches, os.path.join(dst, f"{R}_{probe}_full.csv"))
        c = counters_of(os.path.join(src, f"{R}_{probe}_plain.log"))
        common = lambda l: {"kernel": kernel_name(l), "ncu_ms": num(l, "gpu__time_duration.sum"),  # noqa: E731
                            "fp32_flop": fp32_flop(l),
                            "dram_bytes": num(l, "dram__bytes_read.sum") + num(l, "dram__bytes_write.sum"),
                            "pipe_fma_cycles_active_pct": num(l, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", None),
                            "issue_active_pct": num(l, "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
                            "warps_active_pct": num(l, "sm__warps_active.avg.pct_of_peak_sustained_active", None),
                            "registers": num(l, "launch__registers_per_thread", None)}
        src_note = f"profiles/{R}_{probe}_full.csv (ncu --set full + sass op counters, one launch) / COUNTERS of tools/kernel_probe.py {probe}"
        if probe in ("scan", "scan14"):
            l = launches[0]
            e = common(l)
            e.update(counters=c, source=src_note, dram_bytes_per_launch=e["dram_bytes"],
                     fp32_flop_per_tile_visit=e["fp32_flop"] / max(c.get("tile_visits", 0), 1))
            out["knn_scan" if probe == "scan" else "knn_scan_14d"] = e
        elif probe == "edge":
            ad = next((l for l in launches if "adaptive" in kernel_name(l)), None)
            hd = next((l for l in launches if "hard" in kernel_name(l)), None)
            e = {"counters": c, "source": src_note}
            if ad:
                e["adaptive"] = common(ad)
                e["fp32_flop_per_eval_adaptive"] = e["adaptive"]["fp32_flop"] / max(c.get("evals_adaptive", 0), 1)
            if hd:
                e["hard"] = common(hd)
                e["fp32_flop_per_eval_hard"] = e["hard"]["fp32_flop"] / max(c.get("evals_hard", 0), 1)
            both = [x for x in (ad, hd) if x]
            e["dram_bytes_per_launch_pair"] = sum(num(l, "dram__bytes_read.sum") + num(l, "dram__bytes_write.sum") for l in both)
            t = sum(num(l, "gpu__time_duration.sum") for l in both)
            e["pipe_fma_cycles_active_pct"] = sum(
                num(l, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") * num(l, "gpu__time_duration.sum")
                for l in both) / max(t, 1e-9)
            e["issue_active_pct"] = sum(
                num(l, "smsp__issue_active.avg.pct_of_peak_sustained_active") * num(l, "gpu__time_duration.sum")
                for l in both) / max(t, 1e-9)
            out["edge_pair"] = e
        else:
            e = {"launches": [common(l) for l in launches], "counters": c, "source": src_note}
            units = c.get("configs")
            if units and launches:
                e["fp32_flop_per_config"] = fp32_flop(launches[0]) / units
            out[{"configs": "collide_configs", "fk": "fk_chain", "accept": "accept_pass"}[probe]] = e
    path = os.path.join(dst, f"{R}_kernel_flops.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path)
    for k, v in out.items():
        print(k, {a: b for a, b in v.items() if not isinstance(b, (dict, list))})


if __name__ == "__main__":
    main()
