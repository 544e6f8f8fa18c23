set -x
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py tests/test_lapl_cube_gpu.py -m gpu -q -x 2>&1 | tail -15
bash scripts/gpu_ab.sh r02c cube1023 "FDMB_RING=1" "FDMB_RING=0" "FDMB_RING=1 FDMB_BLOCKED=1"
