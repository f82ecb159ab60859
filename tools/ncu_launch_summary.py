"""Aggregate an ncu launch list (ncu --metrics gpu__time_duration.sum[,smsp__thread_inst_executed_per_inst_executed.ratio,
sm__inst_executed.sum] --csv --log-file X) per kernel for the LAST frame in the log (from the last primary k_trace launch on).
    python tools/ncu_launch_summary.py launches.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, idi = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
per = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    d = per.setdefault(r[idi], {'k': r[ki]})
    v = float(r[vi].replace(',', ''))
    if r[mi].startswith('gpu__time'):
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1e-6)
        d['ms'] = v
    elif r[mi].startswith('smsp__thread'):
        d['lanes'] = v
    else:
        d['inst'] = v
ids = list(per.keys())
prim = [i for i in ids if 'k_trace<0, 1>' in per[i]['k'] or 'k_trace<1, 1>' in per[i]['k']]
start = ids.index(prim[-1]) if prim else 0
agg = collections.OrderedDict()
for i in ids[start:]:
    d = per[i]
    a = agg.setdefault(d['k'][:40], [0, 0, 0, 0])
    a[0] += d.get('ms', 0); a[1] += 1; a[2] += d.get('inst', 0); a[3] += d.get('inst', 0) * d.get('lanes', 0)
tot = sum(a[0] for a in agg.values())
print('last frame: %d launches, kernel time %.3f ms (cold-cache, serialised: compare shares)' % (sum(a[1] for a in agg.values()), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print('%-42s %8.3f ms %4d launches %5.1f%%  lanes %4.1f  Minst %.0f' % (k, a[0], a[1], 100 * a[0] / tot, a[3] / max(a[2], 1), a[2] / 1e6))
