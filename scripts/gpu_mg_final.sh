# Usage (under gpurun --gpus N): bash scripts/gpu_mg_final.sh N <tag> -- every sharded suite + the default bench line on N GPUs
N=${1:-2}; tag=${2:-r02mg}
mkdir -p gpurun_out/$tag
timeout 1500 python -m pytest tests/test_lapl_cube_sharded_gpu.py tests/test_ns_cube_sharded_gpu.py tests/test_lapl_cyl_sharded_gpu.py tests/test_ns_cyl_sharded_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/$tag/sharded_tests_${N}gpu.txt; tail -3 gpurun_out/$tag/sharded_tests_${N}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N > gpurun_out/$tag/bench_cube1023_${N}gpu.json 2> gpurun_out/$tag/bench_cube1023_${N}gpu.err || tail -5 gpurun_out/$tag/bench_cube1023_${N}gpu.err
python scripts/show_bench.py gpurun_out/$tag/bench_cube1023_${N}gpu.json | grep -v "^    cube"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/$tag/bench_cube1023_${N}gpu.json") if l.startswith("{")][-1])
print({k:{kk:v.get(kk) for kk in ("value","ms_per_step","step_frac")} for k,v in d.get("extra",{}).items()})
PY
