# one full ncu capture of the fused FGH + divergence kernel at 255^3
mkdir -p gpurun_out/r02fgh
B="--no-cpu-baseline --no-e2e-batch --no-extra --no-check"
FDMB_GRAPH=0 FDMB_FGH_FUSED=${1:-3} ncu --set full --clock-control none --import-source on -k regex:k_fgh_div -s 3 -c 1 -o /tmp/fgh -f \
    python bench.py --workload nscube255 --steps 2 --warmup 3 $B > gpurun_out/r02fgh/full_fgh.log 2>&1
ncu -i /tmp/fgh.ncu-rep --page raw --csv > gpurun_out/r02fgh/full_fgh.raw.csv 2>/dev/null
ncu -i /tmp/fgh.ncu-rep --page details > gpurun_out/r02fgh/full_fgh.details.txt 2>/dev/null
ncu -i /tmp/fgh.ncu-rep --page source --csv > gpurun_out/r02fgh/full_fgh.src.csv 2>/dev/null
tail -3 gpurun_out/r02fgh/full_fgh.log
