#!/usr/bin/env python
"""Hot SASS lines of one launch from `ncu -i REP --page source --csv --print-source sass` output (file argument)."""
import csv
import sys

allrows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
print("launches in file:", [allrows[i][1][:50] for i in starts[:-1]])
rows = allrows[starts[which]:starts[which + 1]]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tot = sum(int(r[ix['# Samples']]) for r in data)
totinst = sum(int(r[ix['Instructions Executed']]) for r in data)
print('kernel', rows[0][1][:80] if rows[0] else '', 'samples', tot, 'warp inst', totinst, 'sass lines', len(data))
print("-- top by excessive shared wavefronts")
for r in sorted(data, key=lambda r: -int(r[ix['L1 Wavefronts Shared Excessive']]))[:6]:
    print(r[ix['Source']].strip()[:70].ljust(70), r[ix['L1 Wavefronts Shared Excessive']], r[ix['L1 Wavefronts Shared']], r[ix['Instructions Executed']])
print("-- top by samples")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = {h: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not' not in h and int(r[ix[h]]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(r[ix['Source']].strip()[:70].ljust(70), r[ix['# Samples']], r[ix['Instructions Executed']], top)
agg = {}
for r in data:
    for h in hdr:
        if h.startswith('stall_') and 'Not' not in h:
            agg[h] = agg.get(h, 0) + int(r[ix[h]])
print("-- stall totals", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
