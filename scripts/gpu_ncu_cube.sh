# Usage (under gpurun): bash scripts/gpu_ncu_cube.sh <tag> [workloads]  -- ncu --set full of the 5 sweeps of one solve
tag=${1:-exp}
mkdir -p gpurun_out/$tag
for w in ${2:-cube1023 cube255}; do
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe' -s 15 -c 5 -o gpurun_out/$tag/full_$w -f \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-batch > gpurun_out/$tag/full_$w.log 2>&1
tail -2 gpurun_out/$tag/full_$w.log | cut -c1-300
done
