#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_prof_cube.sh <tag> [workload]
# one `ncu --set full` capture of the five sweeps of one solve + the SASS source pages of x fwd, y fwd and z
tag=$1; w=${2:-cube1023}
out=gpurun_out/$tag; mkdir -p $out
rep=/tmp/ncu_$tag; mkdir -p $rep
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe|k_cols_ring|k_rows_ring' -s 15 -c 5 -o $rep/full_$w -f \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-batch --no-extra --no-check > $out/full_$w.log 2>&1
ncu -i $rep/full_$w.ncu-rep --page raw --csv > $out/full_$w.raw.csv 2>/dev/null
for k in 0 1 2; do
  ncu -i $rep/full_$w.ncu-rep --page source --csv --launch-skip $k --launch-count 1 > $out/src_${w}_$k.csv 2>/dev/null
done
ls -la $out
