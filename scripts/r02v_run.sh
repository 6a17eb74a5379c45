#!/bin/bash
# round-2 session v: full GPU test suite after the backward re-ordering (critical GEMM first, column sums on a second
# auxiliary stream), bench A/B on the main-stream priority, other configurations
mkdir -p gpurun_out
export B200VQA_NO_BUILD=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02v_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02v_pytest_gpu.log
tail -4 gpurun_out/r02v_pytest_gpu.log
for variant in "high" "normal"; do
  timeout 300 python bench.py --no-cpu-baseline --main-priority $variant > gpurun_out/r02v_bench_n1_$variant.json 2>> gpurun_out/r02v_bench.err
done
for cfg in 3 4 5 6; do
  timeout 300 python bench.py --no-cpu-baseline --config $cfg > gpurun_out/r02v_bench_cfg$cfg.json 2>> gpurun_out/r02v_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches_per_step"], round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/r02v_bench.err
