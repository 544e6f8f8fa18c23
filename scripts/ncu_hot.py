"""Hot SASS instructions (by stall samples) of one captured launch.  Usage: ncu_hot.py file.ncu-rep launch_index [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
if rep.endswith(".csv"):      # a source page exported on the GPU box (scripts/gpu_profile.sh); kid is ignored
    out = open(rep, errors="ignore").read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kid), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ci['# Samples']].isdigit()]
S = lambda r, h: int(r[ci[h]] or 0)
tot = sum(S(r, '# Samples') for r in data)
print(rows[0][1][:110]); print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(S(r, h) for r in data) for h in stalls}
print([(k, round(100 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
for i in sorted(range(len(data)), key=lambda i: -S(data[i], '# Samples'))[:topn]:
    r = data[i]
    top = sorted(((h, S(r, h)) for h in stalls), key=lambda kv: -kv[1])[:3]
    print(i, r[ci['Source']][:60], S(r, '# Samples'), top)
    for j in range(max(0, i - 3), min(len(data), i + 2)):
        print('      ', j, data[j][ci['Source']][:70], S(data[j], '# Samples'))
