#!/bin/bash
# SASS census of the shipped library per kernel: tcgen05 (UTCHMMA / UTCBAR / LDTM), TMA (UTMALDG / UBLKCP), legacy
# tensor-core MMAs (HMMA: the router), mbarrier waits (SYNCS).   usage: scripts/sass_census.sh [lib] > profiles/rNN_sass_census.txt
lib=${1:-vqa_model_builder_b200/libb200vqa.so}
echo "# cuobjdump -sass $lib (sm_100a): instruction counts per kernel"
cuobjdump -sass "$lib" | awk '
  /Function :/ { name=$3; next }
  /UTCHMMA/ {u[name]++} /UTCBAR/ {b[name]++} /LDTM/ {l[name]++} /UTMALDG/ {t[name]++} /UBLKCP/ {k[name]++}
  /[ .]HMMA\./ {h[name]++} /SYNCS/ {s[name]++}
  END { for (n in s) names[n]=1; for (n in u) names[n]=1; for (n in h) names[n]=1; for (n in k) names[n]=1;
        printf "%-9s %-7s %-6s %-8s %-7s %-6s %-6s %s\n", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "SYNCS", "kernel";
        for (n in names) printf "%-9d %-7d %-6d %-8d %-7d %-6d %-6d %s\n", u[n], b[n], l[n], t[n], k[n], h[n], s[n], n }' | (read -r hdr; echo "$hdr"; sort -k8) | c++filt | sed -e "s/b200::(anonymous namespace):://" -e "s/(CUtensorMap.*//" -e "s/(.*//"
