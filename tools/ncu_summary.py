"""Summarise an ncu launch-list CSV (gpu__time_duration.sum) per kernel."""
import csv, collections, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg.setdefault(row['Kernel Name'][:70], []).append(v)
for k, v in agg.items():
    print("{:70s} n={:3d} mean={:8.1f}us min={:8.1f} max={:8.1f}".format(k, len(v), sum(v) / len(v), min(v), max(v)))
