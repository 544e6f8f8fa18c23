# merged init_bound kernel: NSCube suites (single + float), then nscube31 / nscube255
timeout 900 python -m pytest tests/test_ns_cube_gpu.py tests/test_f32_gpu.py tests/test_cxx_shim_gpu.py tests/test_velocity_plot_gpu.py -m gpu -q -x 2>&1 | grep -E "passed|failed|Error|error|FAILED|Fatal|test_" | tail -8
for w in nscube31 nscube255; do
  bash scripts/gpu_ab.sh r02bound_$w $w "FDMB_PDL=1" 2>&1 | grep -E "==|steps/s|ns_" | sed 's/ e2e=.*//'
done
