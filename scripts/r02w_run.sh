#!/bin/bash
# round-2 session w: GEMM epilogue operand prefetch + single-launch plan: tests, timeline, micro-benchmarks, bench lines
tag=${1:-r02w}
mkdir -p gpurun_out
export B200VQA_NO_BUILD=1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
B200VQA_GEMM_TRACE=1 timeout 300 python scripts/gemm_trace.py > gpurun_out/${tag}_gemm_trace.txt 2>&1
cat gpurun_out/${tag}_gemm_trace.txt
timeout 300 python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_microbench.txt 2>&1
head -22 gpurun_out/${tag}_gemm_microbench.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --no-cpu-baseline --config 5 > gpurun_out/${tag}_bench_cfg5.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches_per_step"], round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 python scripts/attn_bench.py > gpurun_out/${tag}_attn_bench.txt 2>&1
B200VQA_ATTN=simt timeout 300 python scripts/attn_bench.py >> gpurun_out/${tag}_attn_bench.txt 2>&1
cat gpurun_out/${tag}_attn_bench.txt
