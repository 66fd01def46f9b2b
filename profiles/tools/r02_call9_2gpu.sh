#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "iterate_host or real_coefficient or annihilat" > gpurun_out/r02i_tests.log 2>&1
tail -5 gpurun_out/r02i_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --no-secondary > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err
echo "2gpu rc=$?"; wc -c gpurun_out/r02i_bench_2gpu.json; tail -20 gpurun_out/r02i_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 2 --no-secondary --no-e2e --exchange nccl > gpurun_out/r02i_bench_2gpu_nccl.json 2> gpurun_out/r02i_bench_2gpu_nccl.err
echo "2gpu nccl rc=$?"
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r02i_bench_1gpu.json 2> gpurun_out/r02i_bench_1gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02i_bench*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), d["e2e"].get("h2d_bytes_per_step"), d["e2e"].get("d2h_bytes_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
