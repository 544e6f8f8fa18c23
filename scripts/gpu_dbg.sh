for d in ${2:-0 6}; do
  echo "== FDMB_DBG=$d"; FDMB_DBG=$d bash scripts/gpu_quick.sh "${1:-cube1023}" 5 2>&1 | grep -E "cube_y_fwd|cube_z|cube_x_fwd"
done
