#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_profile_r02.sh <tag>
# ncu evidence of round 2: launch lists (gpu__time_duration) of the bench commands + one `ncu --set full` capture per
# workload; scripts/summarize_ncu.py <tag> turns gpurun_out/<tag>/ into profiles/<tag>_*.md and profiles/traffic.json.
tag=${1:-r02}
out=gpurun_out/$tag; mkdir -p $out
rep=/tmp/ncu_$tag; mkdir -p $rep
B="--no-cpu-baseline --no-e2e-batch --no-extra --no-check"
for w in cube1023 cube255 cube127 nscube255 nscyl128 cyl128; do
  FDMB_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 $B > $out/launches_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe|k_cols_ring|k_rows_ring' -s 15 -c 5 -o $rep/full_cube1023 -f \
    python bench.py --workload cube1023 --steps 2 --warmup 3 $B > $out/full_cube1023.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe' -s 15 -c 5 -o $rep/full_cube255 -f \
    python bench.py --workload cube255 --steps 2 --warmup 3 $B > $out/full_cube255.log 2>&1
FDMB_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'k_fgh|k_rhs|k_update|k_bound' -s 9 -c 3 -o $rep/full_nscube255 -f \
    python bench.py --workload nscube255 --steps 2 --warmup 3 $B > $out/full_nscube255.log 2>&1
FDMB_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'k_cyl|k_tridiag|k_cols|k_rows' -s 33 -c 11 -o $rep/full_nscyl128 -f \
    python bench.py --workload nscyl128 --steps 2 --warmup 3 $B > $out/full_nscyl128.log 2>&1
for w in cube1023 cube255 nscube255 nscyl128; do
  ncu -i $rep/full_$w.ncu-rep --page raw --csv > $out/full_$w.raw.csv 2>/dev/null
done
for k in 0 1 2; do
  ncu -i $rep/full_cube1023.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $out/src_cube1023_$k.csv 2>/dev/null
done
ls -la $out | head -40
