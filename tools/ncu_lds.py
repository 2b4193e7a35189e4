"""Shared-memory wavefronts per LDS/STS opcode (and the worst instructions) from an `ncu --page source --csv` dump."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0])
per = []
tot_inst = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']].strip()
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith('@') else (tok[0] if tok else '')
    n = int(r[ix['Instructions Executed']])
    tot_inst += n
    if op.startswith(('LDS', 'STS')):
        a = agg[op]
        w, i = int(r[ix['L1 Wavefronts Shared']]), int(r[ix['L1 Wavefronts Shared Ideal']])
        a[0] += n; a[1] += w; a[2] += i
        if n:
            per.append((w / n, n, src))
print("instructions", tot_inst)
for k, v in sorted(agg.items()):
    print(k, v, 'wf/inst %.2f ideal/inst %.2f' % (v[1] / max(v[0], 1), v[2] / max(v[0], 1)))
hist = defaultdict(int)
for w, n, s in per:
    hist[(s.split()[0] if not s.startswith('@') else s.split()[1], round(w, 1))] += n
for k in sorted(hist):
    print(k, hist[k])
