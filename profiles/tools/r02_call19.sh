#!/bin/bash
# 1 GPU: full suite on the current build; k_walk / k_singles at 5 CTAs per SM (variants); default command timing
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02s
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
for v in "" _w5 _s5 _w5s5; do
  NECI_GPU_LIB=neci_stable_b200/libneci_gpu${v}.so timeout 300 python bench.py $B > gpurun_out/${T}_bench${v}.json 2> gpurun_out/${T}_bench${v}.err
done
SECONDS=0
timeout 900 python bench.py > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err
echo "default command wall ${SECONDS}s"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02s_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"), d.get("clocks"))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench.err gpurun_out/${T}_default.err
