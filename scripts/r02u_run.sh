#!/bin/bash
# round-2 session u: q-tiled tensor-core attention tests, prefetch A/B, GEMM phase timeline
mkdir -p gpurun_out
export B200VQA_NO_BUILD=1
timeout 900 python -m pytest tests/test_gpu_attn.py tests/test_gpu_dropout.py tests/test_gpu_edge.py tests/test_gpu_decoder.py tests/test_gpu_fusion.py -x -q -m gpu > gpurun_out/r02u_pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02u_pytest_attn.log
tail -5 gpurun_out/r02u_pytest_attn.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_bench_n1.err
timeout 300 python bench.py --no-cpu-baseline --no-prefetch > gpurun_out/r02u_bench_n1_noprefetch.json 2>> gpurun_out/r02u_bench_n1.err
python - <<'PY'
import json
for f in ("r02u_bench_n1.json", "r02u_bench_n1_noprefetch.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
B200VQA_GEMM_TRACE=1 timeout 300 python scripts/gemm_trace.py > gpurun_out/r02u_gemm_trace.txt 2>&1
cat gpurun_out/r02u_gemm_trace.txt
