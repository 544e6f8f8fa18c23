# multi-GPU: sharded suites + bench at N ranks.  bash scripts/gpu_r02l.sh <tag> <N>
tag=$1; N=$2
mkdir -p gpurun_out/$tag
nvidia-smi -L | head -8
timeout 2400 python -m pytest tests/test_lapl_cube_sharded_gpu.py tests/test_ns_cube_sharded_gpu.py tests/test_lapl_cyl_sharded_gpu.py tests/test_ns_cyl_sharded_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/$tag/sharded_tests_${N}gpu.txt; tail -6 gpurun_out/$tag/sharded_tests_${N}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/$tag/bench_${N}gpu.json 2> gpurun_out/$tag/bench_${N}gpu.err || tail -20 gpurun_out/$tag/bench_${N}gpu.err
python scripts/show_bench.py gpurun_out/$tag/bench_${N}gpu.json
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/$tag/bench_${N}gpu.json") if l.startswith("{")][-1])
print("extra:", json.dumps(d.get("extra"))[:900])
print("cpu_baseline:", json.dumps(d.get("cpu_baseline"))[:400])
PY
