mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x 2>&1 | tail -5
bash scripts/gpu_ab.sh r02e cube1023 "FDMB_RING=1"
