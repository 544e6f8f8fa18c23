"""Shared-memory wavefront budget of one exported SASS source page, grouped by opcode and wavefronts-per-instruction.
Usage: ncu_wf.py src_page.csv"""
import csv, io, sys, collections
out = open(sys.argv[1], errors="ignore").read()
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ci["# Samples"]].isdigit()]
I = lambda r, h: int(r[ci[h]] or 0)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in data:
    src = r[ci['Source']].split()
    if not src: continue
    op = src[1] if src[0].startswith('@') else src[0]
    ex, wf = I(r, 'Instructions Executed'), I(r, 'L1 Wavefronts Shared')
    if wf == 0 or ex == 0: continue
    key = (op, round(wf / ex, 1))
    a = agg[key]; a[0] += 1; a[1] += ex; a[2] += wf; a[3] += I(r, '# Samples')
tot = sum(a[2] for a in agg.values())
print(f"total smem wavefronts {tot}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{k[0]:10s} wf/inst {k[1]:4.1f}  sass lines {a[0]:5d}  executed {a[1]:12d}  wavefronts {a[2]:12d} ({100 * a[2] / tot:5.1f} %)  samples {a[3]}")
