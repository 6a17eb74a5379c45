#!/bin/bash
# On the GPU box: the single-GPU evidence set of a round (tag = r02s ...): pytest -m gpu log, smoke, bench lines of every
# configuration, the reference (CPU) arm, the saturating batch, the MOE / GEMM micro-benchmarks.
tag=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_n1_default.json 2> gpurun_out/${tag}_bench_n1_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null
for c in 3 4 5 6; do python bench.py --config $c --no-cpu-baseline > gpurun_out/${tag}_bench_cfg${c}.json 2>/dev/null; done
python bench.py --batch 1024 --no-cpu-baseline --detail > gpurun_out/${tag}_bench_saturating_B1024.json 2> gpurun_out/${tag}_gemm_calls_B1024.txt
python scripts/moe_bench.py > gpurun_out/${tag}_moe_cfg5_kernels.txt 2>&1
python scripts/gemm_bench.py > gpurun_out/${tag}_gemm_microbench.txt 2>&1
for f in gpurun_out/${tag}_bench_*.json; do python scripts/bench_table.py $f 2>/dev/null | head -1; done
head -14 gpurun_out/${tag}_moe_cfg5_kernels.txt
