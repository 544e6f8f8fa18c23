# Usage (under gpurun --gpus N): bash scripts/gpu_mg.sh N [workloads]
N=${1:-2}
WL=${2:-"cube255 cube1023"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lapl_cube_sharded_gpu.py -x -q 2>&1 | tail -15
for w in $WL; do
  echo "== $w N=1"
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mg_${w}_1.json 2> gpurun_out/mg_${w}_1.err || tail -5 gpurun_out/mg_${w}_1.err
  python scripts/show_bench.py gpurun_out/mg_${w}_1.json
  for n in 2 4 8; do
    if [ $n -le $N ]; then
      echo "== $w N=$n"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 \
        bench.py --gpus $n --workload $w --steps 20 --warmup 3 > gpurun_out/mg_${w}_$n.json 2> gpurun_out/mg_${w}_$n.err || tail -5 gpurun_out/mg_${w}_$n.err
      python scripts/show_bench.py gpurun_out/mg_${w}_$n.json
    fi
  done
done
