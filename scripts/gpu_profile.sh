#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_profile.sh <round-tag>
# Produces, under gpurun_out/<tag>/: launch lists (ncu gpu__time_duration) for the bench commands and
# one `ncu --set full` capture per workload; scripts/summarize_ncu.py turns them into profiles/*.
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
for w in ${2:-cube1023 cube255 nscube255}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-batch > $out/launches_$w.log 2>&1
done
# the .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): only their raw / source CSV pages come back
rep=/tmp/ncu_$tag
mkdir -p $rep
for w in cube1023 cube255; do
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe' -s 15 -c 5 -o $rep/full_$w -f \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-batch > $out/full_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_fgh|k_rhs|k_update|k_ns' -s 12 -c 4 -o $rep/full_nscube255 -f \
    python bench.py --workload nscube255 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-batch > $out/full_nscube255.log 2>&1
for w in cube1023 cube255 nscube255; do
  ncu -i $rep/full_$w.ncu-rep --page raw --csv > $out/full_$w.raw.csv 2>/dev/null
done
# SASS-level stall / shared-memory wavefront pages of the three distinct sweeps of the 1023^3 solve
for k in 0 1 2; do
  ncu -i $rep/full_cube1023.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $out/src_cube1023_$k.csv 2>/dev/null
done
ls -la $out
