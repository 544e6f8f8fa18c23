# 8 GPUs, short: the sharded NSCube suite (the kernels changed last) and the default bench line
mkdir -p gpurun_out/r02mg8
timeout 300 python -m pytest tests/test_ns_cube_sharded_gpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02mg8/ns_cube_sharded_8gpu.txt; tail -1 gpurun_out/r02mg8/ns_cube_sharded_8gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02mg8/bench_cube1023_8gpu.json 2> gpurun_out/r02mg8/bench_cube1023_8gpu.err || tail -5 gpurun_out/r02mg8/bench_cube1023_8gpu.err
python scripts/show_bench.py gpurun_out/r02mg8/bench_cube1023_8gpu.json | grep -v "^    cube"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02mg8/bench_cube1023_8gpu.json") if l.startswith("{")][-1])
print({k:{kk:v.get(kk) for kk in ("value","ms_per_step","step_frac")} for k,v in d.get("extra",{}).items()})
PY
