#!/bin/bash
# On the GPU box: ncu evidence of a round's final kernels (tag): launch lists of one step of configs 1 and 5 with
# DRAM / tensor-pipe / SM counters, the GEMM DRAM traffic file bench.py reads, --set full captures of the tensor-core
# attention at 328 x 328 tokens (query-tiled forward; backward dQ-tile and dK/dV-tile roles), SASS census.
tag=$1
export B200VQA_NO_BUILD=1
bash scripts/profile_step.sh ${tag}_cfg1 1 > gpurun_out/${tag}_profile_cfg1.log 2>&1
bash scripts/profile_step.sh ${tag}_cfg5 5 > gpurun_out/${tag}_profile_cfg5.log 2>&1
python scripts/update_traffic.py gpurun_out/${tag}_cfg1_launches.csv 1 32 "profiles/${tag}_launches_metrics_step_cfg1.md (ncu, one eager step)"
python scripts/update_traffic.py gpurun_out/${tag}_cfg5_launches.csv 5 128 "profiles/${tag}_launches_metrics_step_cfg5.md (ncu, one eager step)"
cp profiles/gemm_dram_traffic.json gpurun_out/${tag}_gemm_dram_traffic.json
# attention at T = S = 328: launches 16.. of attn_bench belong to that shape (5 shapes x 3 launches before it)
bash scripts/ncu_source.sh ${tag}_ncu_attn_fwd_q328 attn_fwd_tc 16 -- python scripts/attn_bench.py --iters 1 > /dev/null 2>&1
bash scripts/ncu_source.sh ${tag}_ncu_attn_bwd_q328 attn_bwd_tc 16 -- python scripts/attn_bench.py --iters 1 > /dev/null 2>&1
for k in attn_fwd_q328 attn_bwd_q328; do
  python scripts/summarize_ncu_raw.py gpurun_out/${tag}_ncu_${k}_raw.csv gpurun_out/${tag}_ncu_raw_${k}.md "ncu --set full, ${k} (B=32 H=8 T=S=328 d_h=96)" | head -12
  python scripts/summarize_ncu_source.py gpurun_out/${tag}_ncu_${k}_source.csv > gpurun_out/${tag}_ncu_source_${k}.txt 2>/dev/null
  rm -f gpurun_out/${tag}_ncu_${k}_source.csv
done
bash scripts/sass_census.sh > gpurun_out/${tag}_sass_census.txt 2>&1
ls -la gpurun_out | grep ${tag} | head -40
