#!/bin/bash
# 2 GPUs: the whole GPU suite (single- and multi-GPU tests) on the final build
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
T=r02zz
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests_all.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests_all.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_gpu_tests_all.log | head -20
