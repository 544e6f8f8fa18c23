timeout 1500 python -m pytest tests/test_f32_gpu.py tests/test_cxx_shim_gpu.py tests/test_ns_cube_gpu.py tests/test_lapl_cube_gpu.py tests/test_lapl_rect_gpu.py -m gpu -q 2>&1 | tail -25
timeout 300 python scripts/dbg_f32.py
