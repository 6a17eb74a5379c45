#!/bin/bash
# On the GPU box: profile ONE kernel launch with --set full and source correlation, export the details / raw / source
# pages as CSV and drop the report (it embeds the cubin; gpurun_out is capped at 64 MiB).
# usage: scripts/ncu_source.sh <tag> <kernel-regex> <launch-skip> -- <command...>
tag=$1; regex=$2; skip=$3; shift 4
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$regex" -s "$skip" -c 1 \
  -o gpurun_out/$tag -f "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit=$?"
rep=gpurun_out/$tag.ncu-rep
[ -f "$rep" ] || exit 1
ncu -i "$rep" --page details --csv > gpurun_out/${tag}_details.csv 2>/dev/null
ncu -i "$rep" --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i "$rep" --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
rm -f "$rep"
ls -la gpurun_out/${tag}_*
