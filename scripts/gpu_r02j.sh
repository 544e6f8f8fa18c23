mkdir -p gpurun_out/r02j
timeout 1200 python -m pytest tests/test_ns_cube_gpu.py tests/test_cxx_shim_gpu.py tests/test_velocity_plot_gpu.py tests/test_z_drivers_gpu.py -m gpu -q 2>&1 | tail -8
for w in nscube31 nscube255; do
  bash scripts/gpu_ab.sh r02j_$w $w "FDMB_FGH_FUSED=1" "FDMB_FGH_FUSED=0" 2>&1 | grep -v "^    \(cube\|check\)" 
done
