mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py tests/test_lapl_cube_gpu.py -m gpu -q -x 2>&1 | tail -4
bash scripts/gpu_ab.sh r02x cube1023 "FDMB_XINV_INPLACE=1" "FDMB_XINV_INPLACE=0"
