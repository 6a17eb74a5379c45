#!/bin/bash
# Runs every GPU test file in its own process (a trapped kernel poisons the CUDA context) and keeps the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${@:-tests/test_gpu_gemm.py tests/test_gpu_rowops.py tests/test_gpu_attn.py tests/test_gpu_moe.py tests/test_gpu_fusion.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  code=$?
  echo "$name exit=$code $(tail -1 gpurun_out/$name.log)"
  [ $code -ne 0 ] && rc=1
done
exit $rc
