"""Summarise an `ncu --page source --csv` export of a warp-specialised kernel: samples per barrier wait
(SYNCS ... TRYWAIT + the branch after it), per tcgen05 instruction, and totals between markers.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_regions.py src.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
seen, d = set(), []
for r in rows[2:]:
    if len(r) < len(h) - 2 or not r[ix['# Samples']].strip().isdigit() or r[0] in seen:
        continue
    seen.add(r[0])
    d.append(r)
S = lambda r: int(r[ix['# Samples']])
tot = sum(S(r) for r in d)
print('instructions', len(d), 'samples', tot)
waits = collections.OrderedDict()
for i, r in enumerate(d):
    src = r[ix['Source']]
    if 'TRYWAIT' in src:
        key = re.search(r'\[(.*?)\]', src).group(1)
        w = waits.setdefault(key, [0, int(r[ix['Instructions Executed']]), i])
        w[0] += S(r) + S(d[i + 1])
for k, v in waits.items():
    print(f'wait [{k:24s}] samples {v[0]:8d} ({100 * v[0] / tot:5.1f}%)  first-try executions {v[1]:10d}  @{v[2]}')
marks = []
for i, r in enumerate(d):
    src = r[ix['Source']]
    if any(k in src for k in ('UTCHMMA', 'UTCBAR', 'LDTM', 'UBLKCP', 'EXIT', 'BAR.SYNC', 'STG')):
        marks.append(i)
prev = 0
for i in marks:
    seg = sum(S(r) for r in d[prev:i + 1])
    if seg > tot * 0.002:
        print(f'..@{i:5d} {seg:8d} ({100 * seg / tot:5.1f}%) up to {d[i][ix["Source"]].strip()[:70]}  ex={d[i][ix["Instructions Executed"]]}')
    prev = i + 1
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, b):
        r = d[i]
        print(f"{i:5d} s={S(r):6d} ex={r[ix['Instructions Executed']]:>9} thr={r[ix['Avg. Threads Executed']]:>4} {r[ix['Source']][:90]}")
