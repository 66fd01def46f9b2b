#!/bin/bash
# Everything round 2 needs first, in ONE gpurun call (about 8 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/tools/round2_first_call.sh'
# 1. the whole GPU test suite including the cases gated by NECI_GPU_UNVERIFIED (written after round 1's budget ran out)
# 2. default bench line; the semi-stochastic workload (BASELINE configs[3]) with host-built and device-built rows
# 3. launch list + one full ncu capture of k_determ_spmv (K3) on that workload
# Numbers printed under ncu are never bench values; the bench lines come from the separate runs of step 2.
mkdir -p gpurun_out
W="--workload semistoch_20e40o_pchb --no-cpu-baseline"
NECI_GPU_UNVERIFIED=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1
tail -3 gpurun_out/r02_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python bench.py $W > gpurun_out/r02_bench_semistoch.json 2> gpurun_out/r02_bench_semistoch.err
timeout 300 python bench.py $W --core-build device > gpurun_out/r02_bench_semistoch_devbuild.json 2> gpurun_out/r02_bench_semistoch_devbuild.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_semistoch.csv \
    python bench.py $W --no-e2e --steps 3 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_determ_spmv -s 3 -c 1 -f -o gpurun_out/r02_k3_full \
    python bench.py $W --no-e2e --steps 3 --warmup 3 > gpurun_out/r02_ncu_k3.log 2>&1
# 4. launch-bound variants of the spawning kernel, if they were built in the container beforehand
#    (python -c "from neci_stable_b200 import _build; _build.build_gpu_variant('ctas5', ['K1_CTAS_PER_SM=5'])")
for v in ctas5 ctas6; do
    if [ -f neci_stable_b200/libneci_gpu_$v.so ]; then
        NECI_GPU_LIB=$PWD/neci_stable_b200/libneci_gpu_$v.so timeout 300 python bench.py --no-e2e --no-cpu-baseline \
            > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err
    fi
done
for f in gpurun_out/r02_bench*.json; do echo "== $f"; head -c 400 "$f"; echo; done
