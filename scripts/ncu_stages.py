"""Per-stage budget of one captured sweep kernel: the SASS source page (ncu --page source --csv) cut at its barriers,
instructions / fp64 / shared-memory loads, stores and wavefronts PER TILE.  The page lists the kernel's SASS twice
(source view + disassembly view); only the first copy is counted.
Usage: ncu_stages.py src_page.csv ntiles"""
import csv, io, sys
rows = list(csv.reader(io.StringIO(open(sys.argv[1], errors="ignore").read())))
tiles = float(sys.argv[2])
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ci['# Samples']].isdigit()]
addr = [r[ci['Address']] for r in data]
if addr[0] in addr[1:]:
    data = data[:addr.index(addr[0], 1)]


def I(r, h):
    try:
        return int(r[ci[h]] or 0)
    except ValueError:
        return 0


print(rows[0][1][:100])
print("| segment (ends at) | instr | fp64 | LDS | STS | smem wavefronts | stall samples |")
print("|---|---:|---:|---:|---:|---:|---:|")
seg = dict(n=0, fp=0, lds=0, sts=0, wf=0, smp=0); tot = dict(seg); k = 0
for r in data:
    src = r[ci['Source']]; ex = I(r, 'Instructions Executed'); t = src.split()
    op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else ''))
    seg['n'] += ex; seg['smp'] += I(r, '# Samples'); seg['wf'] += I(r, 'L1 Wavefronts Shared')
    if op.startswith('LDS'): seg['lds'] += ex
    if op.startswith('STS'): seg['sts'] += ex
    if op.split('.')[0] in ('DFMA', 'DADD', 'DMUL'): seg['fp'] += ex
    if (op.startswith('BAR') and ex > 0) or r is data[-1]:
        if seg['n'] / tiles >= 20:
            k += 1
            print(f"| {k}: {src.strip()[:34] if op.startswith('BAR') else 'end'} | {seg['n'] / tiles:.0f} | {seg['fp'] / tiles:.0f} | "
                  f"{seg['lds'] / tiles:.0f} | {seg['sts'] / tiles:.0f} | {seg['wf'] / tiles:.0f} | {seg['smp']} |")
            for q in tot: tot[q] += seg[q]
            seg = dict(n=0, fp=0, lds=0, sts=0, wf=0, smp=0)
print(f"| total | {tot['n'] / tiles:.0f} | {tot['fp'] / tiles:.0f} | {tot['lds'] / tiles:.0f} | {tot['sts'] / tiles:.0f} | {tot['wf'] / tiles:.0f} | {tot['smp']} |")
