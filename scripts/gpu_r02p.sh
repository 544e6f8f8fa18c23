# N-GPU A/B of sharded-solve switches: bash scripts/gpu_r02p.sh <N> "<ENV..>" "<ENV..>" ...
N=$1; shift
mkdir -p gpurun_out/r02p
timeout 900 python -m pytest tests/test_lapl_cube_sharded_gpu.py tests/test_ns_cube_sharded_gpu.py -m gpu -q -x 2>&1 | tail -4
i=0
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620+i)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02p/ab_$i.json 2> gpurun_out/r02p/ab_$i.err || tail -5 gpurun_out/r02p/ab_$i.err
  python scripts/show_bench.py gpurun_out/r02p/ab_$i.json | grep -v "cube_\(x_inv\|y_inv\|mg\)"
  i=$((i+1))
done
