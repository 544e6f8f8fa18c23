# Usage (under gpurun): bash scripts/gpu_check.sh  -- full GPU tests, smoke, quick bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
bash scripts/gpu_quick.sh "${1:-cube127 cube255 cube511 cube1023 nscube255 nscyl128}" ${2:-20}
