"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [top]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    n, t = re.sub(r'\(.*', '', r[4])[:90], float(r[-1]) / 1000
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f'total {tot:.1f} us over {len(rows)} launches')
print('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|')
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f'| `{n}` | {c} | {t:.0f} | {100 * t / tot:.1f}% | {t / c:.1f} |')
