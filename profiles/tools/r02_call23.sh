#!/bin/bash
# 1 GPU: k_generate_w (warp-private tiles) against the shipped k_generate; also the host-latency cuts
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02v
NECI_GPU_LIB=neci_stable_b200/libneci_gpu_genw.so timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests_genw.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests_genw.log
B="--no-e2e --no-cpu-baseline --no-secondary"
for v in "" _genw; do
  NECI_GPU_LIB=neci_stable_b200/libneci_gpu${v}.so timeout 300 python bench.py $B > gpurun_out/${T}_bench${v}.json 2> gpurun_out/${T}_bench${v}.err
  NECI_GPU_LIB=neci_stable_b200/libneci_gpu${v}.so timeout 300 python bench.py $B --workload hubk_6x6 --steps 8 > gpurun_out/${T}_hubk${v}.json 2> gpurun_out/${T}_hubk${v}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench.err gpurun_out/${T}_bench_genw.err
