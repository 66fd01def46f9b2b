#!/bin/bash
# K3 column-blocked SpMV (two load flavours), K1 with inline-alias PCHB entries / 3-word QE records / tile->segment table
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02j
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/${T}_semistoch.json 2> gpurun_out/${T}_semistoch.err
NECI_GPU_LIB=neci_stable_b200/libneci_gpu_spmvvol.so timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/${T}_semistoch_vol.json 2> gpurun_out/${T}_semistoch_vol.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_semistoch.csv \
    python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launches_semistoch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_determ_spmv_blocked|k_determ_finish|k_core_gather" -s 6 -c 3 -f -o gpurun_out/${T}_k3_full \
    python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 3 --warmup 3 > gpurun_out/${T}_ncu_k3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_generate|k_evaluate|k_singles|k_compress|k_annihilate|k_insert|k_list_stats" -s 27 -c 9 -f -o gpurun_out/${T}_k1k2_full \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/${T}_ncu_k1.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02j_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck", {}).get("population_conserved"))
        for k, v in r.get("kernels", {}).items(): print("    ", k, "ms %.4f frac %.3f" % (v["ms_per_launch"], v["frac"]), v.get("csr12_equivalent_gbs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/${T}_bench.err gpurun_out/${T}_semistoch.err
