"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [title]"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
if len(sys.argv) > 2: print(f"# {sys.argv[2]}\n")
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:26]:
    print(f"| `{k[:80]}` | {n} | {t:.0f} | {100*t/tot:.1f}% | {t/n:.1f} |")
print(f"\nTotal {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches.")
