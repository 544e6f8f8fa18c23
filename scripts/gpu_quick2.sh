timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scripts/gpu_quick.sh "cube255 cube1023" 10
