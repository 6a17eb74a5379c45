#!/bin/bash
# On an N-GPU box: multi-GPU correctness check + bench lines (tag, N).
tag=$1; n=$2
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29521 scripts/mgpu_check.py > gpurun_out/${tag}_mgpu_check_${n}gpu.log 2>&1; grep -E "MGPU_CHECK|dp_arena|ep_" gpurun_out/${tag}_mgpu_check_${n}gpu.log | tail -6
run 29522 bench.py --gpus $n > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
for c in ${@:3}; do run $((29530 + c)) bench.py --gpus $n --config $c > gpurun_out/${tag}_bench_cfg${c}_n${n}.json 2> gpurun_out/${tag}_bench_cfg${c}_n${n}.err; done
for f in gpurun_out/${tag}_bench_*n${n}.json; do python scripts/bench_table.py $f 2>/dev/null | head -1; done
