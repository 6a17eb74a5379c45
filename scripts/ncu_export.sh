#!/bin/bash
# On the GPU box: turn every gpurun_out/*.ncu-rep into small CSV pages (details + raw) and drop the 30+ MB report
# (the report embeds the whole cubin; gpurun_out is capped at 64 MiB).
for rep in gpurun_out/*.ncu-rep; do
  [ -f "$rep" ] || continue
  base="${rep%.ncu-rep}"
  ncu -i "$rep" --page details --csv > "${base}_details.csv" 2>/dev/null
  ncu -i "$rep" --page raw --csv > "${base}_raw.csv" 2>/dev/null
  rm -f "$rep"
done
