"""Summarise an ncu launch list (gpu__time_duration.sum csv): per-kernel count, mean, share."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name'][:58]
    agg.setdefault(k, []).append(float(row['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:60s} n={len(v):4d} mean={sum(v)/len(v)/1000:8.2f} us share={sum(v)/tot*100:5.1f}% min={min(v)/1000:.2f} max={max(v)/1000:.2f}")
