#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_profile.sh <round-tag>
# Produces, under gpurun_out/<tag>/: launch lists (ncu gpu__time_duration) for the bench commands and
# one `ncu --set full` capture per workload; scripts/summarize_ncu.py turns them into profiles/*.
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
for w in cube127 cube255 nscube255; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > $out/launches_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 30 -c 5 -o $out/full_cube255 -f \
    python bench.py --workload cube255 --steps 2 --warmup 3 --no-cpu-baseline > $out/full_cube255.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fgh|k_rhs|k_update|k_ns' -s 12 -c 4 -o $out/full_nscube255 -f \
    python bench.py --workload nscube255 --steps 2 --warmup 3 --no-cpu-baseline > $out/full_nscube255.log 2>&1
ls -la $out
