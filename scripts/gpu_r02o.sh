for v in 0 1 2 3; do echo "== FDMB_TRIDIAG=$v"; FDMB_TRIDIAG=$v timeout 300 python bench.py --workload cyl128 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e-batch --no-extra | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step']*1e3, 'us/solve', {k: round(v['ms_per_launch']*1e3,1) for k,v in d['roofline']['kernels'].items()})"; done
