mkdir -p gpurun_out/r02u
timeout 1500 python -m pytest tests/test_lapl_cube_gpu.py tests/test_lapl_cube_large_gpu.py tests/test_lapl_cube_sharded_gpu.py tests/test_ns_cube_sharded_gpu.py -m gpu -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02u/bench_2gpu.json 2> gpurun_out/r02u/bench_2gpu.err || tail -5 gpurun_out/r02u/bench_2gpu.err
python scripts/show_bench.py gpurun_out/r02u/bench_2gpu.json
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e-batch > gpurun_out/r02u/bench_1gpu.json 2> gpurun_out/r02u/bench_1gpu.err; python scripts/show_bench.py gpurun_out/r02u/bench_1gpu.json
