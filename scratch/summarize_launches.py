import csv, collections, io, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
r = csv.DictReader(io.StringIO(''.join(lines)))
agg = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = row['Kernel Name'][:80]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(u, 1)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-82s %5d %10.1f us %5.1f%%' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
print('total', tot / 1e3, 'us')
