mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x -k "kat_device" 2>&1 | tail -40 > gpurun_out/r02f/kat.txt
grep -n "Error\|error" gpurun_out/r02f/kat.txt | head
timeout 900 python -m pytest tests/test_lapl_cube_large_gpu.py -m gpu -q -x -k "kat_device and 1023" 2>&1 | tail -5
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
