#!/bin/bash
# Round 2, 2-GPU call: the multi-rank parity suite (oracle world vs one process per GPU, NCCL and peer-memory exchange,
# load balancing, semi-stochastic) and the weak-scaling bench line at N = 2 with its conservation self-check.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rA > gpurun_out/r02h_multirank_tests.log 2>&1
tail -15 gpurun_out/r02h_multirank_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --no-secondary > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err
timeout 300 python bench.py --no-secondary --no-e2e --no-cpu-baseline > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h_bench*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/r02h_bench_2gpu.err
