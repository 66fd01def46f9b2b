#!/bin/bash
# Round 2, call 3: warp-autonomous K1 + one-block-per-attempt RNG (parity suite first), K1 launch-shape variants,
# TMA-staged K3 against the register-staged one, launch list of the default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02c_gpu_tests.log 2>&1
tail -5 gpurun_out/r02c_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
for v in k1_spt2 k1_c3 k1_b128; do
    NECI_GPU_LIB=$PWD/neci_stable_b200/libneci_gpu_$v.so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-secondary \
        > gpurun_out/r02c_bench_$v.json 2> gpurun_out/r02c_bench_$v.err
done
W="--workload semistoch_20e40o_pchb --no-cpu-baseline --core-build device --steps 10"
timeout 300 python bench.py $W > gpurun_out/r02c_semistoch_default.json 2> gpurun_out/r02c_semistoch_default.err
for v in tma_s4 tma_w8t256 reg_q4c2; do
    NECI_GPU_LIB=$PWD/neci_stable_b200/libneci_gpu_$v.so timeout 300 python bench.py $W \
        > gpurun_out/r02c_semistoch_$v.json 2> gpurun_out/r02c_semistoch_$v.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c_launches.csv \
    python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 3 --warmup 3 > gpurun_out/r02c_ncu_launches.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_bench*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], "K1 frac %.3f" % r["frac"], d.get("selfcheck"))
        if "secondary" in d:
            for k, v in d["secondary"].items(): print("   ", k, json.dumps(v)[:400])
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("gpurun_out/r02c_semistoch_*.json")):
    try:
        d = json.load(open(f)); k = d["roofline"]["kernels"]["k_determ_spmv"]
        print(f, "ms/step %.3f" % d["ms_per_step"], "K3 ms %.4f frac %.3f" % (k["ms_per_launch"], k["frac"]), d["roofline"]["phase_ms_per_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
