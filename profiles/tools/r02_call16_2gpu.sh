#!/bin/bash
# 2 GPUs: multi-rank suite, the driver's default command at N = 2 (device-generated lists, secondary workloads, e2e),
# configs[4] strong scaling point N = 2 (1e9 walkers in total, load balancing), configs[2] shape at 1.25e8 walkers per GPU
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02p
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/${T}_multirank_tests.log 2>&1
tail -3 gpurun_out/${T}_multirank_tests.log
SECONDS=0
timeout 900 $TR --master-port 29517 bench.py --gpus 2 > gpurun_out/${T}_default_2gpu.json 2> gpurun_out/${T}_default_2gpu.err
echo "default N=2 rc=$? wall ${SECONDS}s"; SECONDS=0
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 900 $TR --master-port 29518 bench.py --gpus 2 $B --workload cr2_24e30o_pchb --walkers 1e9 --scaling strong --load-balance --steps 5 --warmup 4 \
    > gpurun_out/${T}_cr2_strong_1e9_2gpu.json 2> gpurun_out/${T}_cr2_strong_1e9_2gpu.err
echo "cr2 strong N=2 rc=$? wall ${SECONDS}s"; SECONDS=0
timeout 900 $TR --master-port 29519 bench.py --gpus 2 $B --workload hubk_6x6 --walkers 1.25e8 --steps 5 --warmup 4 \
    > gpurun_out/${T}_hubk_1p25e8_2gpu.json 2> gpurun_out/${T}_hubk_1p25e8_2gpu.err
echo "hubk N=2 rc=$? wall ${SECONDS}s"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02p_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"), d["config"].get("walkers_total_end"), d["config"].get("load_balance"))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
        for k, v in (d.get("secondary") or {}).items(): print("   sec", k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 4 gpurun_out/${T}_default_2gpu.err gpurun_out/${T}_cr2_strong_1e9_2gpu.err gpurun_out/${T}_hubk_1p25e8_2gpu.err
