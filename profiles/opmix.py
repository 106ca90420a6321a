"""Dynamic SASS opcode mix of one kernel from an ncu report:
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python profiles/opmix.py src.csv"""
import collections
import csv
import re
import sys


def analyze(fn, top=26):
    rows = list(csv.reader(open(fn)))
    h = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[h]
    ia, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    agg = collections.defaultdict(lambda: [0, 0])
    tot = ts = 0
    for r in rows[h + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)', r[ia])
        if not m:
            continue
        op = m.group(2)
        base = 'IMAD.MOV' if op.startswith('IMAD.MOV') else op.split('.')[0]
        n, s = int(r[ie]), int(r[iss] or 0)
        agg[base][0] += n; agg[base][1] += s; tot += n; ts += s
    print(fn, 'warp-instructions', tot, 'samples', ts)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'   {k:12s} {v[0]:12d} {100 * v[0] / tot:5.1f}%   stall samples {100 * v[1] / max(ts, 1):5.1f}%')


if __name__ == '__main__':
    for f in sys.argv[1:]:
        analyze(f)
