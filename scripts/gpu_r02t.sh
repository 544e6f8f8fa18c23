mkdir -p gpurun_out/r02t
timeout 1200 python -m pytest tests/test_lapl_cube_gpu.py tests/test_lapl_cube_large_gpu.py tests/test_ns_cube_gpu.py -m gpu -q 2>&1 | tail -6
for c in 1 4 8 16; do echo "== FDMB_HOST_CHUNKS=$c"; FDMB_HOST_CHUNKS=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-e2e-batch > gpurun_out/r02t/ab_$c.json 2> gpurun_out/r02t/ab_$c.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02t/ab_$c.json") if l.startswith("{")][-1])
print(d["e2e"]["value"], "Gpts/s e2e", d["e2e"]["ms_per_step"], "ms", "device", d["ms_per_step"])
PY
done
