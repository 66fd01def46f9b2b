#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_gpu_tests.log 2>&1
tail -3 gpurun_out/r02d_gpu_tests.log
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
for v in k1_spt2 k1_spt2_c5; do
    NECI_GPU_LIB=$PWD/neci_stable_b200/libneci_gpu_$v.so timeout 300 python bench.py $B > gpurun_out/r02d_bench_$v.json 2> gpurun_out/r02d_bench_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spawn -s 6 -c 1 -f -o gpurun_out/r02d_k1_full \
    python bench.py $B --steps 3 --warmup 3 > gpurun_out/r02d_ncu_k1.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02d_bench*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
