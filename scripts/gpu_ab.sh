# A/B of env switches on one workload: bash scripts/gpu_ab.sh <tag> <workload> "<ENV=.. ENV=..>" "<ENV=..>" ...
tag=$1; wl=$2; shift 2
mkdir -p gpurun_out/$tag
i=0
for envs in "$@"; do
  echo "== $envs"
  env $envs timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-e2e-batch --no-extra > gpurun_out/$tag/ab_$i.json 2> gpurun_out/$tag/ab_$i.err || tail -5 gpurun_out/$tag/ab_$i.err
  python scripts/show_bench.py gpurun_out/$tag/ab_$i.json
  i=$((i+1))
done
