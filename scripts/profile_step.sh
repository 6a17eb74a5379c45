#!/bin/bash
# On the GPU box: ncu launch list (duration, DRAM bytes/throughput, SM throughput, tensor pipe, warps active) of ONE
# eager step of a bench configuration, single stream.   usage: scripts/profile_step.sh <tag> <config> [extra one_step args]
tag=$1; cfg=$2; shift 2
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv \
  python scripts/one_step.py --config $cfg --single-stream "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit=$?"
python scripts/summarize_ncu_metrics.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.md "one step of bench.py --config $cfg $*"
head -45 gpurun_out/${tag}_launches.md | cut -c1-230
