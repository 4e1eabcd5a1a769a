"""The parts of bench.py's contract that can be checked without a GPU: the reference arm prints one JSON line with
the agreed keys (and only rank 0 prints under a multi-rank launch), and the product arm refuses to run without a CUDA
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None, tmp=None):
    e = dict(os.environ, **(env or {}))
    if tmp:
        e["BSG_BENCH_DIR"] = str(tmp)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line(tmp_path):
    r = run_bench("--impl", "reference", "--gscale", "0.002", "--steps", "2", "--warmup", "1", tmp=tmp_path)
    assert r.returncode == 0, r.stderr
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads counted/sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["config"]["reads_in_bam"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["config"]["workload"].startswith("c2: bamProfile") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    assert cb["kind"] == ("reference" if O.ref_available() else "port")      # oracle/_ref = the reference's own engine
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["configs"]["c1"]["cpu_ms_per_call_median"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "int32"


def test_reference_arm_other_ranks_stay_silent(tmp_path):
    r = run_bench("--impl", "reference", "--gpus", "2", "--gscale", "0.002", "--steps", "1", "--warmup", "0",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, tmp=tmp_path)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu(tmp_path):
    import bamsignals_b200 as B
    if B.lib().bsg_device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    r = run_bench("--gscale", "0.002", "--steps", "1", "--warmup", "1", tmp=tmp_path)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
