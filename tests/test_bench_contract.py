"""CPU checks of bench.py's bookkeeping: algorithmic work matches BASELINE.md, the reference arm prints
one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench

    b, l, f = bench.algorithmic_bytes(256, 55, 128, 4, 4)
    assert abs(b / 1e6 - 275.7) < 0.1 and abs(l / 1e6 - 20.44) < 0.01 and abs(f / 1e9 - 25.38) < 0.01
    assert abs((b + 12 * l) / 1e6 - 521.0) < 0.1  # headline unit
    b1, l1, f1 = bench.algorithmic_bytes(128, 46, 62, 4, 3)
    assert abs(b1 / 1e6 - 45.9) < 0.1 and abs(l1 / 1e6 - 5.18) < 0.01 and abs(f1 / 1e9 - 2.08) < 0.01


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "stock", "raft", "core", "corr.py"))
    # the reference's own file when build() has staged it (baseline/_ref), the oracle's port of it otherwise
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]
