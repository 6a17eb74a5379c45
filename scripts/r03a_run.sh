#!/bin/bash
# cluster split-K for small-M GEMMs: guarded tests first (a hung kernel dies with its process), then A/B bench lines
mkdir -p gpurun_out
export B200VQA_NO_BUILD=1
timeout 180 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/r03a_pytest_gemm.log 2>&1
rc=$?; echo "gemm pytest rc=$rc"; tail -15 gpurun_out/r03a_pytest_gemm.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r03a_pytest_gpu.log 2>&1
echo "full pytest rc=$?"; tail -3 gpurun_out/r03a_pytest_gpu.log
for v in 1 0; do
  for c in 1 4; do
    B200VQA_GEMM_CLUSTER=$v timeout 300 python bench.py --no-cpu-baseline --config $c 2>/dev/null > gpurun_out/r03a_bench_cfg${c}_cluster$v.json
    python -c "
import json; d=json.loads(open('gpurun_out/r03a_bench_cfg${c}_cluster$v.json').read().strip().splitlines()[-1]); print('cfg$c cluster=$v', d['ms_per_step'], round(d['value']), round(d['e2e']['value']), d['gpu_launches_per_step'], round(d['roofline']['frac'],4))"
  done
done
timeout 200 python bench.py --no-cpu-baseline --detail 2>&1 >/dev/null | grep "M=    32" | head -12
