#!/bin/bash
# Usage (under gpurun): bash scripts/gpu_prof_cyl.sh <tag>
# ncu evidence for the cylindrical family at BASELINE configs[3] (128 x 127 x 128): launch list + one full capture
# of the solve's five sweeps and of the NS step's stencil kernels.
tag=$1
out=gpurun_out/$tag; mkdir -p $out
rep=/tmp/ncu_$tag; mkdir -p $rep
for w in nscyl128 cyl128; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv \
      python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-batch > $out/launches_$w.log 2>&1
done
FDMB_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'k_cyl|k_tridiag|k_cols|k_rows' -s 33 -c 11 -o $rep/full_nscyl128 -f \
    python bench.py --workload nscyl128 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-batch > $out/full_nscyl128.log 2>&1
ncu -i $rep/full_nscyl128.ncu-rep --page raw --csv > $out/full_nscyl128.raw.csv 2>/dev/null
# SASS page of the tridiagonal kernel (first one in the capture window)
ncu -i $rep/full_nscyl128.ncu-rep --page source --csv -k regex:'k_tridiag' --launch-count 1 > $out/src_tridiag.csv 2>/dev/null
ls -la $out
