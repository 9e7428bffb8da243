"""The JSON-line contract of bench.py, checked without a GPU: the reference arm is run for real (a few frames on the host
cores), the GPU arm through its kept record profiles/evidence_r2/bench_c5.json (written by `python bench.py` on a B200)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _last_json(text):
    return json.loads([ln for ln in text.strip().splitlines() if ln.startswith("{")][-1])


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _last_json(out.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("front-end frames/s") and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port") and cb["cores"] >= 1
    assert cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "c5_zed_streams"
    # both arms describe the workload with the same dict
    gpu = _last_json(open(os.path.join(ROOT, "profiles", "evidence_r2", "bench_c5.json")).read())
    assert gpu["config"] == d["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_kept_gpu_record_has_the_contract_keys():
    d = _last_json(open(os.path.join(ROOT, "profiles", "evidence_r2", "bench_c5.json")).read())
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic" and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] is not None
    # achieved = algorithmic bytes per launch / live launch time
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 64 * 1280 * 720 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not d["clocks"]["reasons"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["parity"]["ids_and_stereo_bits_equal"] is True
    assert d["parity"]["max_px_err_vs_ref_cpu"] <= d["parity"]["tolerance_px"] == 0.02
