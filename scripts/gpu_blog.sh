for bl in 4 6; do for d in 0 6; do
  echo "== BLOG=$bl DBG=$d"; FDMB_BLOG=$bl FDMB_DBG=$d bash scripts/gpu_quick.sh cube1023 5 2>&1 | grep -E "cube_y_fwd|cube_z|cube_x_fwd|cube_x_inv"
done; done
echo "== natural"; FDMB_BLOCKED=0 bash scripts/gpu_quick.sh cube1023 5 2>&1 | grep -E "cube_y_fwd|cube_z|cube_x_fwd|cube_x_inv"
