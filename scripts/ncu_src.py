"""Aggregate an ncu source-page CSV (SASS view) of one kernel: shared wavefronts and stall samples by opcode class.
Usage: ncu_src.py file.ncu-rep kernel_index(0-based among captured)"""
import csv, io, subprocess, sys, collections
rep, kid = sys.argv[1], int(sys.argv[2])
if rep.endswith(".csv"):      # a source page exported on the GPU box (scripts/gpu_profile.sh); kid is ignored
    out = open(rep, errors="ignore").read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kid), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
lines = out.split('"Kernel Name",')[1].splitlines()
print(lines[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for r in rows[1:]:
    if len(r) < len(hdr): continue
    op = r[ci["Source"]].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.rstrip(";")
    key = o
    vals = [int(r[ci["Instructions Executed"]] or 0), int(r[ci["L1 Wavefronts Shared"]] or 0),
            int(r[ci["L1 Wavefronts Shared Ideal"]] or 0), int(r[ci["# Samples"]] or 0),
            int(r[ci["L2 Theoretical Sectors Global"]] or 0)]
    for i, v in enumerate(vals):
        agg[key][i] += v; tot[i] += v
print(f"{'opcode':28s} {'inst':>12s} {'smem_wf':>12s} {'ideal':>12s} {'samples':>8s} {'l2sect':>12s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][3])[:40]:
    print(f"{k:28s} {v[0]:12d} {v[1]:12d} {v[2]:12d} {v[3]:8d} {v[4]:12d}")
print(f"{'TOTAL':28s} {tot[0]:12d} {tot[1]:12d} {tot[2]:12d} {tot[3]:8d} {tot[4]:12d}")
