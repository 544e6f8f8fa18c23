"""Shared-memory wavefronts per SASS instruction of one exported source page (excess over ideal = bank conflicts).
Usage: ncu_conf.py src_page.csv [top]"""
import csv, io, sys
out = open(sys.argv[1], errors="ignore").read()
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ci["# Samples"]].isdigit()]
I = lambda r, h: int(r[ci[h]] or 0)
tot = sum(I(r, 'L1 Wavefronts Shared') for r in data); ideal = sum(I(r, 'L1 Wavefronts Shared Ideal') for r in data)
print(rows[0][1][:100] if len(rows[0]) > 1 else '', 'wavefronts', tot, 'ideal', ideal, 'excess', tot - ideal)
lds = sum(I(r, 'L1 Wavefronts Shared') for r in data if 'LDS' in r[ci['Source']]); sts = sum(I(r, 'L1 Wavefronts Shared') for r in data if 'STS' in r[ci['Source']])
print('LDS wavefronts', lds, 'STS wavefronts', sts, 'other', tot - lds - sts)
ex = sorted(range(len(data)), key=lambda i: -(I(data[i], 'L1 Wavefronts Shared') - I(data[i], 'L1 Wavefronts Shared Ideal')))
for i in ex[:topn]:
    r = data[i]
    print(i, r[ci['Source']][:64], 'exec', I(r, 'Instructions Executed'), 'wf', I(r, 'L1 Wavefronts Shared'), 'ideal', I(r, 'L1 Wavefronts Shared Ideal'))
