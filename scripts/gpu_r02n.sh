mkdir -p gpurun_out/r02n
timeout 1500 python -m pytest tests/test_ns_cyl_gpu.py tests/test_lapl_cyl_gpu.py tests/test_lapl_rect_gpu.py tests/test_velocity_plot_gpu.py tests/test_ns_cube_gpu.py -m gpu -q 2>&1 | tail -8
for w in nscyl128 cyl128; do bash scripts/gpu_ab.sh r02n_$w $w "FDMB_GRAPH=1"; done
