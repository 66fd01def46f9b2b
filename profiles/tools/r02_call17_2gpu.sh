#!/bin/bash
# 2 GPUs: load balancing with a populated reference determinant (configs[4] logic at 2e8 walkers in total)
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02q
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 900 $TR --master-port 29518 bench.py --gpus 2 $B --workload cr2_24e30o_pchb --walkers 2e8 --scaling strong --load-balance --steps 5 --warmup 4 \
    > gpurun_out/${T}_cr2_strong_2e8_2gpu.json 2> gpurun_out/${T}_cr2_strong_2e8_2gpu.err
echo "cr2 strong N=2 rc=$?"
timeout 300 python bench.py $B --workload cr2_24e30o_pchb --walkers 1e8 --ref-fraction 0.05 --steps 5 --warmup 4 > gpurun_out/${T}_cr2_ref_1gpu.json 2> gpurun_out/${T}_cr2_ref_1gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02q_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"), d["config"].get("walkers_total_end"), d["config"].get("load_balance"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 6 gpurun_out/${T}_cr2_strong_2e8_2gpu.err gpurun_out/${T}_cr2_ref_1gpu.err
