mkdir -p gpurun_out/exp
for z in 0 1; do
  if [ $z = 1 ]; then export FDMB_ZEXP=1; fi
  timeout 300 python bench.py --workload cube1023 --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['unit'], 'ms/step', d['ms_per_step'])
for k,v in d['roofline']['kernels'].items(): print('   ',k, round(v['ms_per_launch']*1e3,1),'us')
"
done
unset FDMB_ZEXP
ncu --set full --clock-control none --import-source on -k regex:'k_cols_pipe|k_rows_pipe' -s 15 -c 5 -o gpurun_out/exp/full_cube1023 -f \
    python bench.py --workload cube1023 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/exp/full_cube1023.log 2>&1
tail -3 gpurun_out/exp/full_cube1023.log
timeout 600 python -m pytest tests/test_lapl_rect_gpu.py -x -q 2>&1 | tail -15
