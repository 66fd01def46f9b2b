#!/bin/bash
# Round 2, call 2: the whole (un-gated) GPU suite, K3 launch-shape variants on the semi-stochastic workload, a full
# ncu capture of k_determ_spmv (shipped build) and a source-level capture of k_spawn on the default workload.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02b_gpu_tests.log 2>&1
tail -3 gpurun_out/r02b_gpu_tests.log
W="--workload semistoch_20e40o_pchb --no-cpu-baseline --core-build device --steps 10"
timeout 300 python bench.py $W > gpurun_out/r02b_semistoch_default.json 2> gpurun_out/r02b_semistoch_default.err
for v in q1c8 q2c5 q4c3 q4c2; do
    if [ -f neci_stable_b200/libneci_gpu_$v.so ]; then
        NECI_GPU_LIB=$PWD/neci_stable_b200/libneci_gpu_$v.so timeout 300 python bench.py $W \
            > gpurun_out/r02b_semistoch_$v.json 2> gpurun_out/r02b_semistoch_$v.err
    fi
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_determ_spmv -s 3 -c 1 -f -o gpurun_out/r02b_k3_full \
    python bench.py $W --no-e2e --steps 3 --warmup 3 > gpurun_out/r02b_ncu_k3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spawn -s 6 -c 1 -f -o gpurun_out/r02b_k1_full \
    python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r02b_ncu_k1.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02b_semistoch_*.json")):
    try:
        d = json.load(open(f)); k = d["roofline"]["kernels"]["k_determ_spmv"]
        print(f, "ms/step %.3f" % d["ms_per_step"], "K3 ms %.4f frac %.3f" % (k["ms_per_launch"], k["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
