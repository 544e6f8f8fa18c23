# whole GPU suite, all failures listed: bash scripts/gpu_tests.sh <tag>
mkdir -p gpurun_out/$1
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/$1/gpu_tests.txt; tail -15 gpurun_out/$1/gpu_tests.txt
