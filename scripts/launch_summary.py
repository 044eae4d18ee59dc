"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name'][:64]; v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000.0 if unit == 'ns' else (v * 1000.0 if unit == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:5d} {a[1]:10.1f}us {100*a[1]/tot:5.1f}%  avg {a[1]/a[0]:8.1f} max {a[2]:8.1f}  {k}")
