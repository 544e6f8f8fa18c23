# final validation of a round: whole GPU suite, smoke, reference arm, default bench line
tag=${1:-r02z}
mkdir -p gpurun_out/$tag
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/$tag/gpu_tests_1gpu.txt; tail -4 gpurun_out/$tag/gpu_tests_1gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/$tag/bench_ref.json 2> gpurun_out/$tag/bench_ref.err; tail -c 600 gpurun_out/$tag/bench_ref.json
timeout 900 python bench.py > gpurun_out/$tag/bench_cube1023.json 2> gpurun_out/$tag/bench_cube1023.err; python scripts/show_bench.py gpurun_out/$tag/bench_cube1023.json; tail -3 gpurun_out/$tag/bench_cube1023.err
for w in cube127 cube255 cube511 nscube31 nscube255 nscyl128 cyl128; do timeout 600 python bench.py --workload $w --steps 100 --warmup 5 --no-extra > gpurun_out/$tag/bench_$w.json 2> gpurun_out/$tag/bench_$w.err; python scripts/show_bench.py gpurun_out/$tag/bench_$w.json | head -3; done
