#!/bin/bash
# 2 GPUs: multi-rank suite and the default bench with destination-grouped remote stores and the early clock sampler
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02t
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/${T}_multirank_tests.log 2>&1
tail -3 gpurun_out/${T}_multirank_tests.log
B="--no-e2e --no-cpu-baseline --no-secondary"
NECI_GPU_TIMING=1 timeout 600 $TR --master-port 29517 bench.py --gpus 2 $B > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "rc=$?"; grep "peer-memory" gpurun_out/${T}_bench_2gpu.err
timeout 600 $TR --master-port 29518 bench.py --gpus 2 $B --workload hubk_6x6 --steps 8 > gpurun_out/${T}_hubk_2gpu.json 2> gpurun_out/${T}_hubk_2gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02t_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_bench_2gpu.err
