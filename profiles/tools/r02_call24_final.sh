#!/bin/bash
# final single-GPU verification of the shipped build: suite, the driver's default command, reference arm, launch list,
# full ncu capture of the K1 / K2 kernels, sanitizers on smoke()
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02w
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.log 2>&1
tail -3 gpurun_out/${T}_gpu_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests.log | head -20
SECONDS=0
timeout 900 python bench.py > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err
echo "default command wall ${SECONDS}s"; SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_reference_arm.err
echo "reference arm wall ${SECONDS}s"
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_generate|k_evaluate|k_singles|k_compress|k_annihilate|k_insert|k_list_stats" -s 30 -c 10 -f -o gpurun_out/${T}_k1k2_full \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/${T}_ncu_k1k2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_compress|k_annihilate|k_insert" -s 9 -c 3 -f -o gpurun_out/${T}_k2_cr2_full \
    python bench.py $B --workload cr2_24e30o_pchb --steps 2 --warmup 3 > gpurun_out/${T}_ncu_k2_cr2.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/${T}_sanitizer_racecheck.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.3e ms/step %s" % (d["value"], d.get("ms_per_step")), (d.get("roofline") or {}).get("phase_ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_default.err gpurun_out/${T}_reference_arm.err
