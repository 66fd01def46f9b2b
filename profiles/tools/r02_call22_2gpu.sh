#!/bin/bash
# 2 GPUs: configs[4] strong-scaling point N = 2 with the settings of the N = 4 / 8 points (5 % of the walkers on the reference)
mkdir -p gpurun_out
T=r02u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 900 $TR --master-port 29518 bench.py --gpus 2 $B --workload cr2_24e30o_pchb --walkers 1e9 --scaling strong --load-balance --steps 5 --warmup 4 \
    > gpurun_out/${T}_cr2_strong_1e9_2gpu.json 2> gpurun_out/${T}_cr2_strong_1e9_2gpu.err
echo "rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02u_cr2_strong_1e9_2gpu.json") if l.startswith("{")][-1])
print("%.3e" % d["value"], d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["selfcheck"], d["config"].get("load_balance"))
PY
tail -n 3 gpurun_out/${T}_cr2_strong_1e9_2gpu.err
