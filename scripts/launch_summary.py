"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)   # -> us
    grid = row.get('Grid Size', '')
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1e3:.3f} ms total device time (serialised, cold cache)")
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:80]}` | {n} | {t/1e3:.3f} | {100*t/tot:.1f}% | {t/n:.1f} |")
