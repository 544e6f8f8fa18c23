mkdir -p gpurun_out/r02i
timeout 1200 python -m pytest tests/test_ns_cube_gpu.py tests/test_ns_cyl_gpu.py tests/test_lapl_cube_gpu.py tests/test_cxx_shim_gpu.py tests/test_nbody_gpu.py tests/test_velocity_plot_gpu.py tests/test_z_drivers_gpu.py -m gpu -q 2>&1 | tail -8
for w in nscube31 nscube255 nscyl128 cube127; do
  bash scripts/gpu_ab.sh r02i_$w $w "FDMB_GRAPH=1" "FDMB_GRAPH=0" 2>&1 | grep -v "^    \(cube\|ns\|check\)" 
done
