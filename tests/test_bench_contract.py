"""bench.py's contract, as far as a box without a GPU can check it: the reference arm (the unmodified reference binary,
or the oracle port when it is not built) prints ONE JSON line with the agreed keys on BASELINE config 1; our arm refuses
to run without a CUDA device instead of falling back to the CPU; the committed bench lines of the round carry the
roofline / cpu_baseline / e2e objects."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_json_line(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "0", "--workdir", str(tmp_path)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["metric"] == "query_intervals_per_s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["config"]["workload"].startswith("cfg1")
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the arm would run")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "cfg1", "--steps", "1",
                        "--workdir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout) and "{" not in r.stdout


def test_committed_bench_lines_carry_the_contract_objects():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r3*_bench_cfg3*.json")))
    assert files
    for f in files:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        assert BASE_KEYS <= set(j), f
        assert j["scaling"] == "strong" and j["config"]["workload"].startswith("cfg3"), f
        if "PROFILING AID" in j["config"].get("parallelism", ""):
            continue
        rf, e2e = j["roofline"], j["e2e"]
        assert rf["bound"] == "hbm" and 0 < rf["frac"] <= 1.2 and rf["peak"] > 0 and rf["unit"] == "GB/s", f
        assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 12 * j["config"]["n_queries"], f
        assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 1, f
        assert j["parity_vs_oracle"] is True and j["gpu_launches"] > 0, f
