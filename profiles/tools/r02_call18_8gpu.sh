#!/bin/bash
# 8 GPUs: the driver's default command at N = 8, configs[2] (6x6 k-space Hubbard, 1e8 and 1e9 walkers hashed over 8 GPUs),
# configs[4] (Cr2-sized 24e/30o, 1e9 walkers, strong-scaling point N = 8 with load balancing)
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02u
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus_${N}.txt
B="--no-e2e --no-cpu-baseline --no-secondary"
SECONDS=0
NECI_GPU_TIMING=1 timeout 900 $TR --master-port 29517 bench.py --gpus $N > gpurun_out/${T}_default_${N}gpu.json 2> gpurun_out/${T}_default_${N}gpu.err
echo "default N=$N rc=$? wall ${SECONDS}s"; SECONDS=0
grep "peer-memory" gpurun_out/${T}_default_${N}gpu.err | head -8
timeout 600 $TR --master-port 29518 bench.py --gpus $N $B --workload cr2_24e30o_pchb --walkers 1e9 --scaling strong --load-balance --steps 5 --warmup 4 \
    > gpurun_out/${T}_cr2_strong_1e9_${N}gpu.json 2> gpurun_out/${T}_cr2_strong_1e9_${N}gpu.err
echo "cr2 strong rc=$? wall ${SECONDS}s"; SECONDS=0
if [ "$N" = "8" ]; then
timeout 600 $TR --master-port 29519 bench.py --gpus $N $B --workload hubk_6x6 --walkers 1.25e8 --steps 5 --warmup 4 \
    > gpurun_out/${T}_hubk_1e9_${N}gpu.json 2> gpurun_out/${T}_hubk_1e9_${N}gpu.err
echo "hubk 1e9 rc=$? wall ${SECONDS}s"; SECONDS=0
timeout 600 $TR --master-port 29520 bench.py --gpus $N $B --workload hubk_6x6 --walkers 1.25e7 --steps 10 --warmup 4 \
    > gpurun_out/${T}_hubk_1e8_${N}gpu.json 2> gpurun_out/${T}_hubk_1e8_${N}gpu.err
echo "hubk 1e8 rc=$? wall ${SECONDS}s"
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02u_*gpu.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"], d.get("selfcheck"), d["config"].get("walkers_total_end"), d["config"].get("load_balance"))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
        for k, v in (d.get("secondary") or {}).items(): print("   sec", k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 4 gpurun_out/${T}_*_${N}gpu.err | grep -v "^\*\|OMP_NUM\|^$"
