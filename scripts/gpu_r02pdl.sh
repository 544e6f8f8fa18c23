# programmatic dependent launch on the launch-bound sizes: whole GPU suite, then default / FDMB_PDL=0 A/B
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in nscube31 cube127 nscyl128 cyl128 cube255; do
  bash scripts/gpu_ab.sh r02pdl_$w $w "FDMB_PDL=1" "FDMB_PDL=0" 2>&1 | grep -E "==|steps/s|Gpts/s" | sed 's/ step_frac.*//'
done
