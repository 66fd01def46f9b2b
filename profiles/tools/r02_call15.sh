#!/bin/bash
# 1 GPU: device-generated start lists; the per-GPU shapes of configs[2] and configs[4] (1.25e8 walkers per GPU)
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02o
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "synthetic_list" > gpurun_out/${T}_tests.log 2>&1
tail -5 gpurun_out/${T}_tests.log
B="--no-e2e --no-cpu-baseline --no-secondary"
SECONDS=0; timeout 300 python bench.py $B --list device > gpurun_out/${T}_bench_devlist.json 2> gpurun_out/${T}_bench_devlist.err
echo "dev1e7 wall ${SECONDS}s"; SECONDS=0; timeout 600 python bench.py $B --workload hubk_6x6 --walkers 1.25e8 --steps 5 --warmup 4 > gpurun_out/${T}_hubk_1p25e8.json 2> gpurun_out/${T}_hubk_1p25e8.err
echo "hubk wall ${SECONDS}s"; SECONDS=0; timeout 600 python bench.py $B --workload cr2_24e30o_pchb --walkers 1.25e8 --steps 5 --warmup 4 > gpurun_out/${T}_cr2_1p25e8.json 2> gpurun_out/${T}_cr2_1p25e8.err
echo "cr2 wall ${SECONDS}s"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02o_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"), d["config"].get("walkers_total_end"), d["config"].get("determinants_total_end"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 5 gpurun_out/${T}_bench_devlist.err gpurun_out/${T}_hubk_1p25e8.err gpurun_out/${T}_cr2_1p25e8.err
