echo "== B nosync 1023^3"; timeout 300 python scripts/dbg_kat.py nosync 1023 1023 1023 2>&1 | tail -4
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x 2>&1 | tail -5
bash scripts/gpu_ab.sh r02h cube1023 "FDMB_RING=1"
