mkdir -p gpurun_out/r02k
timeout 1200 python -m pytest tests/test_ns_cube_gpu.py -m gpu -q 2>&1 | tail -4
bash scripts/gpu_ab.sh r02k nscube255 "FDMB_FGH_FUSED=1" 2>&1 | grep -v "^    \(cube\|check\)"
