"""Aggregate an ncu --page source CSV by CUDA source line: instructions executed + stall samples.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [topN]
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows[:60]) if 'Instructions Executed' in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
# in cuda,sass mode each SASS row carries the source line in col 0/1
inst = defaultdict(int)
samp = defaultdict(int)
src = {}
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    ln = r[0]
    if ln:
        cur = ln
        src[cur] = r[1]
    try:
        inst[cur] += int(r[ci['Instructions Executed']] or 0)
        samp[cur] += int(r[ci['# Samples']] or 0)
    except ValueError:
        pass
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print('total warp-instructions %d, samples %d' % (tot_i, tot_s))
for ln, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print('%6s  inst %9d (%4.1f%%)  samples %6d (%4.1f%%)  %s' % (ln, n, 100.0 * n / max(tot_i, 1), samp[ln],
                                                                   100.0 * samp[ln] / max(tot_s, 1), src.get(ln, '')[:110]))
