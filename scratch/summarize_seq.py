"""Print the launch sequence (kernel, grid, duration) of an ncu gpu__time_duration csv, skipping fills."""
import csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
for x in csv.DictReader(io.StringIO(''.join(lines))):
    if x.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = x['Kernel Name'].replace('dl4ds::', '')
    if 'elementwise' in k or 'distribution' in k:
        continue
    v = float(x['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(x['Metric Unit'], 1)
    print('%-70s grid=%-14s %8.1f us' % (k[:70], x['Grid Size'], v))
