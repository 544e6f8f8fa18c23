# Usage (under gpurun --gpus N): bash scripts/gpu_mg_quick.sh N "workloads" [steps]
N=${1:-2}
for w in ${2:-cube1023}; do
  echo "== $w N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --workload $w --steps ${3:-10} --warmup 3 --no-cpu-baseline > gpurun_out/mgq_${w}_$N.json 2> gpurun_out/mgq_${w}_$N.err || tail -5 gpurun_out/mgq_${w}_$N.err
  python scripts/show_bench.py gpurun_out/mgq_${w}_$N.json
done
