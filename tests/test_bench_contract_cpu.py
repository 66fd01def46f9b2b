"""bench.py contract on CPU: the reference arm (CPU restatement on the host cores) prints one JSON line with the keys
the driver reads, for the default workload and for the semi-stochastic one; the engine arm refuses to run without a
CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    return p


@pytest.mark.parametrize("extra", [[], ["--workload", "semistoch_20e40o_pchb", "--core-size", "1500", "--trial", "2"]])
def test_reference_arm_prints_the_contract_line(extra):
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "3", "--cpu-sample-walkers", "4e4", *extra)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spawn_attempts_per_sec" and d["unit"] == "attempts/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == (extra[1] if extra else "n2_14e28o_pchb")
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"


def test_engine_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-e2e")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
