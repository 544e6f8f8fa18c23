echo "== A sync 1023^3"; timeout 300 python scripts/dbg_kat.py sync 1023 1023 1023 2>&1 | tail -4
echo "== B nosync 1023^3"; timeout 300 python scripts/dbg_kat.py nosync 1023 1023 1023 2>&1 | tail -4
echo "== C nosync 1023^3 RING=0"; FDMB_RING=0 timeout 300 python scripts/dbg_kat.py nosync 1023 1023 1023 2>&1 | tail -4
echo "== D sanitizer nosync 15x1023x1023"; timeout 600 compute-sanitizer --tool memcheck python scripts/dbg_kat.py nosync 15 1023 1023 2>&1 | tail -25
