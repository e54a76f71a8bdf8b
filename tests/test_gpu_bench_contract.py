"""GPU: bench.py prints ONE JSON line carrying every key of the measurement contract, for the headline
workload (with the CPU baseline / parity leg) and for a side workload."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import ROOT  # noqa: E402


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_headline_line_has_every_contract_key():
    d = _run("--pool", "832", "--steps", "2", "--warmup", "3", "--e2e-seconds", "0.3", "--leg-seconds", "0.3", "--eer-n", "3000000",
             "--cpu-seconds", "3")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "utterances/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "l2" in d["config"]
    assert d["value"] > 1e4 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 1e3 and e["h2d_bytes_per_step"] == 832 * 321 * 180 * 4 and e["d2h_bytes_per_step"] == 832 * 4
    assert e["value"] != d["value"] and e["steps"] >= 10 and 0 < e["frac_of_h2d_ceiling"] < 1.5 and e["h2d_ceiling_gbs"] > 1
    # the other BASELINE configs ride in the same line, each with its own roofline and a clock record taken under its load
    w = d["workloads"]
    assert set(w) == {"cae", "hybrid", "cnn1d", "eer"}
    for name, leg in w.items():
        assert leg["value"] > 0 and leg["roofline"]["frac"] > 0 and leg["clocks"]["samples"] >= 3, (name, leg["clocks"])
    assert w["cae"]["roofline"]["bound"] == "tensor" and w["cnn1d"]["roofline"]["bound"] == "hbm" and w["eer"]["roofline"]["bound"] == "hbm"
    h = w["hybrid"]["e2e"]
    assert h["scores_identical_to_separate_calls"] is True and h["h2d_bytes_per_step"] == 832 * 321 * 180 * 4
    assert h["three_uploads"]["h2d_bytes_per_step"] == 3 * h["h2d_bytes_per_step"]
    assert w["eer"]["eer_select"]["identical_result"] is True and set(w["eer"]["inputs"]) == {"affine", "permuted", "sigmoid"}
    assert all(v["identical_result"] for v in w["eer"]["inputs"].values())
    r = d["roofline"]
    assert r["frac_of_burst_peak"] < r["frac"] and r["frac_of_nominal_peak"] < r["frac_of_burst_peak"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["parity"]["max_rel_err_scores_vs_cpu_reference"] <= d["parity"]["tolerance"] == 1e-3
    assert d["e2e_f16_slab"]["scores_identical_to_fp32_slab"] is True
    f32 = d["parity"]["fp32_mode"]                      # the same utterances through the full-fp32 kernels (precision="fp32")
    assert f32["max_rel_err_scores_vs_cpu_reference"] <= 5e-6 and f32["utterances_per_s"] > 0 and "eer_delta_pp" in f32
    sp = d["parity"]["split_mode"]                      # ... and through the split-precision tensor-core kernels (precision="split")
    assert sp["max_rel_err_scores_vs_cpu_reference"] <= 5e-6 and sp["utterances_per_s"] > 10 * f32["utterances_per_s"]


def test_eer_workload_reports_both_paths():
    d = _run("--workload", "eer", "--eer-n", "3000000", "--leg-seconds", "0.3")
    assert d["unit"] == "scores/s" and d["roofline"]["bound"] == "hbm" and d["scaling"] == "replicas only"
    assert d["eer_select"]["identical_result"] is True and d["eer_select"]["value"] > d["value"]
