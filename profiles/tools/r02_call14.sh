#!/bin/bash
# 1 GPU: direct opposite-spin pair pick in the k-space generator, restored K3 kernel, full suite
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02n
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/${T}_semistoch.json 2> gpurun_out/${T}_semistoch.err
timeout 300 python bench.py $B --workload hubk_6x6 --steps 8 > gpurun_out/${T}_hubk.json 2> gpurun_out/${T}_hubk.err
timeout 300 python bench.py $B --workload cr2_24e30o_pchb --steps 8 > gpurun_out/${T}_cr2.json 2> gpurun_out/${T}_cr2.err
timeout 300 python bench.py $B --workload hubrs_4x4 --steps 8 > gpurun_out/${T}_hubrs.json 2> gpurun_out/${T}_hubrs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_hubk.csv \
    python bench.py $B --workload hubk_6x6 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launches_hubk.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02n_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
        for k, v in r.get("kernels", {}).items(): print("    ", k, "ms %.4f frac %.3f" % (v["ms_per_launch"], v["frac"]), v.get("csr12_equivalent_gbs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench.err gpurun_out/${T}_hubk.err gpurun_out/${T}_hubrs.err
