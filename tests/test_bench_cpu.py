"""bench.py contract pieces that need no GPU: the reference arm (the oracle port on the host
cores -- one of the places allowed to execute oracle/), the ncu traffic reader, and the rule
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


def test_ncu_traffic_reader():
    sys.path.insert(0, ROOT)
    import bench
    t = bench.ncu_traffic(os.path.join(ROOT, "profiles", "r01_knn_scan_full.csv"))
    assert isinstance(t, int) and 1e8 < t < 1e10          # ~0.4 GB per launch: the packed rows stay in L2
    assert bench.ncu_traffic(os.path.join(ROOT, "profiles", "no_such_file.csv")) is None


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--nodes",
                        "2000", "--no-cpu"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and '"value"' not in r.stdout
