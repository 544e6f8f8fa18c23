tag=${1:-r01b}
out=gpurun_out/$tag
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe' -s 20 -c 5 -o $out/full_cube255 -f \
    python bench.py --workload cube255 --steps 2 --warmup 3 --no-cpu-baseline > $out/full_cube255.log 2>&1
tail -3 $out/full_cube255.log
