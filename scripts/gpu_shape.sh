for d in 0 1 3; do
for shp in "1023 7 1023" "1023 63 1023" "1023 255 1023" "1023 1023 255"; do
  echo "== DBG=$d shape $shp"; FDMB_DBG=$d python scripts/prof_shape.py $shp 3 2>&1 | grep -E "cube_[yz]_fwd|Error|error"
done; done
