# 2-GPU A/B of the sharded overlap: bash scripts/gpu_r02p.sh
mkdir -p gpurun_out/r02p
timeout 900 python -m pytest tests/test_lapl_cube_sharded_gpu.py -m gpu -q -x 2>&1 | tail -4
i=0
for envs in "FDMB_MG_OVERLAP=0" "FDMB_MG_OVERLAP=4" "FDMB_MG_OVERLAP=4 FDMB_MG_SPLIT=56" "FDMB_MG_OVERLAP=4 FDMB_MG_SPLIT=92" "FDMB_MG_OVERLAP=8"; do
  echo "== $envs"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29620+i)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02p/ab_$i.json 2> gpurun_out/r02p/ab_$i.err || tail -5 gpurun_out/r02p/ab_$i.err
  python scripts/show_bench.py gpurun_out/r02p/ab_$i.json | grep -v "cube_\(x_inv\|y_inv\|mg\)"
  i=$((i+1))
done
