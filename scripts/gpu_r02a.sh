# round-2 first GPU call: the whole GPU suite, then the default bench line and the reference arm
set -x
mkdir -p gpurun_out/r02a
nvidia-smi -L; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02a/gpu_tests.txt; tail -5 gpurun_out/r02a/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err; tail -c 1500 gpurun_out/r02a/bench_ref.json
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02a/bench_cube1023.json 2> gpurun_out/r02a/bench_cube1023.err; tail -c 3000 gpurun_out/r02a/bench_cube1023.json; tail -5 gpurun_out/r02a/bench_cube1023.err
