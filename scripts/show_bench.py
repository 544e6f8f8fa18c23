"""Pretty-print the JSON line(s) bench.py wrote to a file."""
import json
import sys

for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline", {})
    print(f"{d.get('value'):.4g} {d.get('unit')} n_gpus={d.get('n_gpus')} ms/step={d.get('ms_per_step'):.4g} "
          f"step_frac={r.get('step_frac', 0):.3f} e2e={d.get('e2e', {}).get('value', 0):.4g} clocks={d.get('clocks')}")
    if "check" in d:
        print(f"    check rel_l2 = {d['check']['rel_l2']:.3e} ok={d['check']['ok']}")
    if "sharded" in r:
        sh = r["sharded"]
        print(f"    aggregate HBM+NVLink roofline: {sh['t_roof_ms_no_overlap']:.3f} ms no-overlap -> frac {sh['frac_no_overlap']:.3f}")
    for k, v in r.get("kernels", {}).items():
        print(f"    {k:28s} {v['ms_per_launch'] * 1e3:9.1f} us x{v['launches_per_step']:g}")
