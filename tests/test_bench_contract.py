"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the
CPU restatement of the reference on the host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-scale", "0.02"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mpt_nodes_keccak_hashed_per_sec" and d["unit"] == "nodes/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_defaults_name_the_headline_config():
    """The default run is config 2 of BASELINE.json on one GPU with K/W that finish within minutes."""
    sys.path.insert(0, ROOT)
    import bench

    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--gpus", type=int, default=1)' in src
    assert 'add_argument("--steps", type=int, default=5)' in src and 'add_argument("--warmup", type=int, default=3)' in src
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "20k touched accounts" in baseline["configs"][1] and bench.C2_PARAMS  # the generator parameters of that config
