#!/bin/bash
# 1 GPU: pipelined K3 (U = 1/2/3), k_generate at 5 CTAs/SM, hubk profile, 1e8 walkers
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02m
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for v in "" _spmvu1 _spmvu3; do
  NECI_GPU_LIB=neci_stable_b200/libneci_gpu${v}.so timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/${T}_semistoch${v}.json 2> gpurun_out/${T}_semistoch${v}.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_determ_spmv_blocked" -s 2 -c 1 -f -o gpurun_out/${T}_k3_full \
    python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 3 --warmup 3 > gpurun_out/${T}_ncu_k3.log 2>&1
timeout 300 python bench.py $B --workload hubk_6x6 --steps 8 > gpurun_out/${T}_hubk.json 2> gpurun_out/${T}_hubk.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_generate|k_evaluate" -s 12 -c 4 -f -o gpurun_out/${T}_hubk_k1_full \
    python bench.py $B --workload hubk_6x6 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_hubk.log 2>&1
timeout 400 python bench.py $B --walkers 1e8 --steps 8 > gpurun_out/${T}_bench_1e8.json 2> gpurun_out/${T}_bench_1e8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_1e8.csv \
    python bench.py $B --walkers 1e8 --steps 2 --warmup 3 > gpurun_out/${T}_ncu_launches_1e8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_compress|k_annihilate|k_insert" -s 9 -c 3 -f -o gpurun_out/${T}_k2_1e8_full \
    python bench.py $B --walkers 1e8 --steps 2 --warmup 3 > gpurun_out/${T}_ncu_k2_1e8.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02m_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
        for k, v in r.get("kernels", {}).items(): print("    ", k, "ms %.4f frac %.3f" % (v["ms_per_launch"], v["frac"]), v.get("csr12_equivalent_gbs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench.err gpurun_out/${T}_semistoch.err gpurun_out/${T}_bench_1e8.err
