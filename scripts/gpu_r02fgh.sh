# register-marching fused FGH + divergence (k_fgh_div) variants against k_fgh + k_rhs at 255^3; the whole NSCube suite with it
FDMB_FGH_FUSED=1 timeout 900 python -m pytest tests/test_ns_cube_gpu.py -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_ab.sh r02fgh nscube255 "FDMB_FGH_FUSED=0" "FDMB_FGH_FUSED=1" "FDMB_FGH_FUSED=2" "FDMB_FGH_FUSED=3" "FDMB_FGH_FUSED=5" 2>&1 | grep -v "^    \(cube\|check\|ns_bound\)"
