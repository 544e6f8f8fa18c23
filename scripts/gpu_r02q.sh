# 8-GPU: sharded suites (log to commit) + overlap A/B + the default line with extras
N=${1:-8}
mkdir -p gpurun_out/r02q
nvidia-smi -L | wc -l
timeout 2400 python -m pytest tests/test_lapl_cube_sharded_gpu.py tests/test_ns_cube_sharded_gpu.py tests/test_lapl_cyl_sharded_gpu.py tests/test_ns_cyl_sharded_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02q/sharded_tests_${N}gpu.txt; tail -4 gpurun_out/r02q/sharded_tests_${N}gpu.txt
i=0
for envs in "FDMB_MG_OVERLAP=0" "FDMB_MG_OVERLAP=8" "FDMB_MG_OVERLAP=4 FDMB_MG_SPLIT=40"; do
  echo "== $envs"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29640+i)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02q/ab_$i.json 2> gpurun_out/r02q/ab_$i.err || tail -5 gpurun_out/r02q/ab_$i.err
  python scripts/show_bench.py gpurun_out/r02q/ab_$i.json
  i=$((i+1))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02q/bench_${N}gpu.json 2> gpurun_out/r02q/bench_${N}gpu.err || tail -5 gpurun_out/r02q/bench_${N}gpu.err
python scripts/show_bench.py gpurun_out/r02q/bench_${N}gpu.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02q/bench_${N}gpu.json") if l.startswith("{")][-1])
print("extra:", json.dumps(d.get("extra"))[:1200])
PY
