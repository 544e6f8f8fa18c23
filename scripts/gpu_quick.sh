# quick perf check: bash scripts/gpu_quick.sh "<workloads>" [steps]
for w in ${1:-cube127 cube255}; do echo "== $w"; timeout 300 python bench.py --workload $w --steps ${2:-50} --warmup 5 --no-cpu-baseline --no-e2e-batch | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'step_frac', d['roofline']['step_frac'], 'e2e', d['e2e']['value'])
for k,v in d['roofline']['kernels'].items(): print('   ',k, round(v['ms_per_launch']*1e3,1),'us')
"; done
