"""bench.py prints ONE JSON line with the contract's keys (driver contract + this tier's roofline / cpu_baseline)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

gpu = pytest.mark.gpu

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"}


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "bench.py must print exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


@gpu
def test_bench_line_contract_sc():
    d = _run(["--workload", "sc256", "--steps", "3", "--warmup", "3", "--batch", "65536"])
    assert BASE_KEYS <= set(d) and "cpu_baseline" in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["config"]["workload"] == "sc256"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["gpu_launches"] > 0 and d["value"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_bench_reference_arm_contract():
    d = _run(["--impl", "reference", "--workload", "sc64", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


@gpu
def test_bench_parity_checked_and_identical_config():
    """The product line carries the in-bench oracle check of a timed step's output, and both arms print the same
    `config` object."""
    d = _run(["--workload", "sc256", "--steps", "3", "--warmup", "3", "--batch", "65536", "--no-cpu-baseline"])
    p = d["parity_checked"]
    assert p["ok"] is True and p["rows"] >= 1000 and p["mismatching_rows"] == 0
    r = _run(["--impl", "reference", "--workload", "sc256", "--steps", "1", "--warmup", "0", "--batch", "65536"])
    assert {k: d["config"][k] for k in r["config"]} == r["config"]


@gpu
def test_bench_fused_sweep_workload():
    d = _run(["--workload", "mc256", "--steps", "3", "--warmup", "3", "--batch", "1048576", "--no-cpu-baseline"])
    assert d["frames"] == 3 * 1048576 and 0 < d["bler"] < 1 and d["parity_checked"]["ok"] is True
    assert d["parity_checked"]["counts"] == d["parity_checked"]["oracle_counts"]
    assert d["e2e"]["d2h_bytes_per_step"] == 24 and d["e2e"]["value"] > 0


def test_reference_arm_never_imports_the_product():
    src = open(os.path.join(ROOT, "oracle", "cpu_arm.py")).read()
    assert "import neural_polar_decoder_b200" not in src and "from neural_polar_decoder_b200" not in src
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--workload','sc64','--steps','1','--warmup','0'];"
            "import runpy\ntry:\n runpy.run_path(%r, run_name='__main__')\nexcept SystemExit: pass\n"
            "assert not any(m.startswith('neural_polar_decoder_b200') for m in sys.modules), 'product imported'") % os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
