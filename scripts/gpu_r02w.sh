out=gpurun_out/r02w; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_aux.csv python scripts/prof_aux.py > $out/launches_aux.log 2>&1; tail -2 $out/launches_aux.log
ncu --set full --clock-control none -k regex:'k_tridiag_rows|k_pm_|k_vplot' -c 14 -o /tmp/full_aux -f python scripts/prof_aux.py > $out/full_aux.log 2>&1
ncu -i /tmp/full_aux.ncu-rep --page raw --csv > $out/full_aux.raw.csv 2>/dev/null
ls -la $out
