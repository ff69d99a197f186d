"""CPU-side checks of bench.py: the measured-peak lookup and the reference arm's JSON line
(`bench.py --impl reference` times the reference's own OpenMP path; it needs no GPU)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_measured_peak_lookup():
    b = _bench()
    assert b.find_hbm_peak({})[0] == 6650.0 and "fallback" in b.find_hbm_peak({})[1]
    assert b.find_hbm_peak({"hbm_gbs": 6451.8, "bf16_tflops": 1404.0, "nvlink_bandwidth_gbs": 900.0}) == \
        (6451.8, "MEASURED_PEAKS.json hbm_gbs")
    v, src = b.find_hbm_peak({"hbm": {"burst_gbs": 7000.0, "sustained_gbs": 6451.8}, "bf16": {"sustained_tflops": 1404.0}})
    assert v == 7000.0 and src.endswith("hbm.burst_gbs")              # kernels timed alone: the burst figure
    assert b.find_hbm_peak({"copy_bytes_per_s": 6.4518e12})[0] == pytest.approx(6451.8)
    assert b.find_hbm_peak({"l2_bandwidth_gbs": 9000.0, "pcie_h2d_gbs": 50.0})[0] == 6650.0


def test_reference_arm_prints_one_json_line():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "plmdca_grad_evals_per_sec" and d["unit"] == "evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["N"] == 5000 and d["config"]["L"] == 150 and d["config"]["q"] == 21
