#!/bin/bash
mkdir -p gpurun_out; TAG=${TAG:-r02c3b}
for c in cfg3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_${c}_launches.csv python scratch/prof_cfg34.py $c > gpurun_out/${TAG}_${c}_ncu.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_${c}_launches.csv > gpurun_out/${TAG}_${c}_launches_summary.txt 2>&1
echo "== $c"; head -14 gpurun_out/${TAG}_${c}_launches_summary.txt; tail -1 gpurun_out/${TAG}_${c}_launches_summary.txt
done
python - <<'PY'
import csv
f='gpurun_out/r02c3b_cfg3_launches.csv'
lines=[l for l in open(f) if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
def us(r):
    v=float(r['Metric Value'].replace(',','')); return v/1000 if r['Metric Unit']=='ns' else v
big=sorted(rows,key=lambda r:-us(r))[:12]
for r in big: print('%8.1f us %s grid %s block %s'%(us(r), r['Kernel Name'][:70], r['Grid Size'], r['Block Size']))
PY
