mkdir -p gpurun_out/r02x
bash scripts/gpu_ab.sh r02x cube1023 "FDMB_RING_PAIR=0" "FDMB_RING_PAIR=2" "FDMB_RING_PAIR=3" "FDMB_RING_PAIR=0"
timeout 900 env FDMB_RING_PAIR=3 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x 2>&1 | tail -3
