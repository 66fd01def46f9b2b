#!/bin/bash
# 1 GPU: PCHB particle selection FULL-FULL (parity with the oracle), the whole suite again after the refactoring of
# gen_pchb_double, default bench
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02z
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py $B --particle-selection FULL-FULL > gpurun_out/${T}_bench_fullfull.json 2> gpurun_out/${T}_bench_fullfull.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02z_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench.err gpurun_out/${T}_bench_fullfull.err
