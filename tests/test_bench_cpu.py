"""bench.py contract pieces that need no GPU: the reference arm (the oracle port on the host
cores -- one of the places allowed to execute oracle/), the profile readers, the arm layout, and the rule
that the product arm fails loudly without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
        "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-nodes", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "collision_checked_edges_per_s" and line["unit"] == "edges/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_kernel_flops_arithmetic_and_missing_profile():
    """tools/kernel_flops.py: the FP32 work of a launch is 2*(FFMA + 2*FFMA2) + FADD + 2*FADD2 + FMUL + 2*FMUL2
    thread-level instructions; bench.py reads the per-unit figures from profiles/ and degrades to
    null fractions, not to invented ones, when the file is missing."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench
    import kernel_flops
    launch = {f"sm__sass_thread_inst_executed_op_{n}_pred_on.sum": (str(v), "inst")
              for n, v in (("ffma", 100), ("ffma2", 10), ("fadd", 7), ("fadd2", 1), ("fmul", 5), ("fmul2", 2))}
    assert kernel_flops.fp32_flop(launch) == 2 * (100 + 20) + 7 + 2 + 5 + 4
    assert kernel_flops.num({"dram__bytes_read.sum": ("1.5", "Mbyte")}, "dram__bytes_read.sum") == 1.5e6
    assert bench.load_json(os.path.join(ROOT, "profiles", "no_such_file.json"), {}) == {}


def test_arm_layout_and_csv_row():
    sys.path.insert(0, ROOT)
    import bench
    from mmmp_b200.roadmap import arm_layout
    assert arm_layout(1, 4) == [[0]] * 4
    assert arm_layout(2, 4) == [[0], [1], [0], [1]]
    assert arm_layout(4, 4) == [[0], [1], [2], [3]]
    assert arm_layout(8, 4) == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert arm_layout(6, 4) == [list(range(6))] * 4          # no even split: every arm over all ranks
    for ws in (1, 2, 3, 4, 8, 16):                           # every rank has work, every arm an owner
        lay = arm_layout(ws, 4)
        assert set(r for ranks in lay for r in ranks) == set(range(ws))
    class Teams:
        ranks_of_arm = arm_layout(4, 2)
    lim = bench.limiter([{"knn_scan": 10.0, "edge_collision": 8.0, "accept_pass": 1.0},
                         {"knn_scan": 12.0, "edge_collision": 8.5, "accept_pass": 1.0, "gather_adjacency": 2.0},
                         {"knn_scan": 9.0, "edge_collision": 8.0, "accept_pass": 1.0},
                         {"knn_scan": 9.5, "edge_collision": 8.0, "accept_pass": 1.0}], Teams)
    assert lim["slowest_rank"] == 1 and lim["its_largest_phases"][0] == ("knn_scan", 12.0) and lim["team_size"] == 2
    assert lim["ms_not_shrinking_with_the_team"] == 3.0 and lim["spread_between_ranks_ms"] == 5.5
    row = bench.csv_row(0.25, 1000, 9000, 0.6)
    assert row == {"learning_time": 0.25, "n_nodes": 1000, "n_edges": 4500, "avg_degree": 9.0, "max_edge_len": 0.6}


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--nodes",
                        "2000", "--no-cpu"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and '"value"' not in r.stdout
