"""bench.py contract checks that run without a GPU: the reference arm prints ONE JSON line with the keys the driver reads
(a tiny shape here: the arm itself is the reference's own CPU classes), and the argument defaults are the BASELINE workload."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--box", "64", "--mlr", "16", "--phases", "2",
                          "--mreco", "10", "--cpu-sample", "2", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["metric"] == base["metric"]
    for key in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_defaults_are_the_baseline_workload():
    """with no flags: N = 1, the 100k x 256^2 x (125 x 16) x 9 workload, K / W that finish within minutes"""
    sys.path.insert(0, str(ROOT))
    import importlib
    bench = importlib.import_module("bench")
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert (a.gpus, a.box, a.particles, a.mlr, a.phases, a.mlt, a.mreco) == (1, 256, 100000, 125, 16, 9, 100)
    assert a.warmup >= 3 and 1 <= a.steps <= 10
    wl = bench.workload(a)
    assert wl["r"] == 127 and wl["pf"] == 2
    cfg = bench.config_dict(a, wl, 25134, 25135, 1)
    assert cfg["workload"].startswith("100k synthetic particles, box 256, 2000 orientation samples")


def test_reference_arm_of_a_global_search_iteration():
    """--scan-nr: the reference arm runs the reference's own scan loop (ref_scan), its post-scan Particle logic and the phases on the
    same flags as our arm (tiny shape; the defaults leave the global scan off)"""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not present")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--box", "64", "--mlr", "16", "--phases", "2", "--mreco", "10",
                          "--scan-nr", "120", "--nt", "12", "--rmax", "10", "--cpu-sample", "3", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-600:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["value"] > 0
    assert "global_search" in d["config"] and "120 shared rotations x 12 translations" in d["config"]["global_search"]
    assert d["config"]["r"] == 10
    sys.path.insert(0, str(ROOT))
    import importlib
    bench = importlib.import_module("bench")
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.scan_nr == 0 and a.rmax == 0 and "global_search" not in bench.config_dict(a, bench.workload(a), 25134, 25135, 1)
