#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02g_gpu_tests.log 2>&1
tail -3 gpurun_out/r02g_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/r02g_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02g_launches.csv \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/r02g_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_generate|k_evaluate|k_singles" -s 15 -c 5 -f -o gpurun_out/r02g_k1_full \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/r02g_ncu_k1.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/r02g_semistoch.json 2> gpurun_out/r02g_semistoch.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-secondary --workload hubk_6x6 --steps 8 > gpurun_out/r02g_hubk.json 2> gpurun_out/r02g_hubk.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/r02g_bench.err
