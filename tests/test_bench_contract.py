"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints exactly one JSON line with the keys the
driver reads; the GPU arm's line is checked by the driver itself at round end (profiles/r1_bench_final.json is its shape)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--N", "9", "--Nsub", "20",
                          "--cpu-seeds", "2", "--steps", "1", "--warmup", "1", *extra], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run()
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "starship_flip PTR" in d["config"]["workload"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--N", "9",
                          "--Nsub", "20", "--cpu-seeds", "2", "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_committed_gpu_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "e2e", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["warmup"] >= 3
    k1 = d["roofline_k1"]
    assert k1["bound"] == "fp64" and k1["unit"] == "TFLOP/s" and abs(k1["frac"] - k1["achieved"] / k1["peak"]) < 1e-12
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["clocks"]["reasons"] == []
    assert set(d["config"]) == {"workload", "batch_total", "batch_per_gpu", "partition", "algorithm_constants", "seeds", "l2"}
