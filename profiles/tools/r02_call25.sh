#!/bin/bash
# 1 GPU: K3 with the bank-aware element order inside the chunks
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02x
timeout 900 python -m pytest tests -m gpu -q -x -k "semi or determ or core or energies" > gpurun_out/${T}_k3_tests.log 2>&1
tail -3 gpurun_out/${T}_k3_tests.log
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 8 > gpurun_out/${T}_semistoch.json 2> gpurun_out/${T}_semistoch.err
timeout 300 python bench.py $B --workload semistoch_20e40o_pchb --core-build host --steps 8 > gpurun_out/${T}_semistoch_hostrows.json 2> gpurun_out/${T}_semistoch_hostrows.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_determ_spmv_blocked" -s 2 -c 1 -f -o gpurun_out/${T}_k3_full \
    python bench.py $B --workload semistoch_20e40o_pchb --core-build device --steps 3 --warmup 3 > gpurun_out/${T}_ncu_k3.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02x_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1]); r = d["roofline"]
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), r["phase_ms_per_step"])
        for k, v in r.get("kernels", {}).items(): print("    ", k, "ms %.4f frac %.3f" % (v["ms_per_launch"], v["frac"]), v.get("csr12_equivalent_gbs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/${T}_semistoch.err gpurun_out/${T}_semistoch_hostrows.err
