# Usage (under gpurun): bash scripts/gpu_shape.sh -- per-kernel times of the same sweeps on differently shaped grids
# (the access-pattern table of profiles/r01g_access_pattern.md was made with this and a temporary copy-only flag)
for shp in "1023 7 1023" "1023 63 1023" "1023 255 1023" "1023 1023 255"; do
  echo "== shape $shp"; python scripts/prof_shape.py $shp 3 2>&1 | grep -E "cube_[yz]_|Error|error"
done
