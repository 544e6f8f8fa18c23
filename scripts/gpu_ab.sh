set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 compute-sanitizer --tool memcheck python -c "
import numpy as np, fdm_b200
from oracle import fdm_oracle as O
n=63; dx=1.0/n; l=1+dx
rhs=O.synthetic_rhs((n,n,n),seed=3)
a=fdm_b200.LaplCube(dx,dx,dx,l,l,l,n,n,n).solve(rhs)
print('err', O.rel_l2(a, O.LaplCube(dx,dx,dx,l,l,l,n,n,n).solve(rhs)))
" 2>&1 | tail -8
for w in cube127 cube255 cube511; do for p in 1 0; do echo "== $w pipe=$p"; FDMB_PIPE=$p timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-e2e-batch | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'step_frac', d['roofline']['step_frac'])
for k,v in d['roofline']['kernels'].items(): print('   ',k, round(v['ms_per_launch']*1e3,1),'us')
"; done; done
