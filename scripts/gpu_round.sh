set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
for w in cube127 cube255 cube511 nscube31 nscube255; do timeout 600 python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 1800 gpurun_out/bench_$w.json; done
