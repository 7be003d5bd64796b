#!/bin/bash
mkdir -p gpurun_out; TAG=${TAG:-r02c34}
for c in cfg3 cfg4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_${c}_launches.csv python scratch/prof_cfg34.py $c > gpurun_out/${TAG}_${c}_ncu.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_${c}_launches.csv > gpurun_out/${TAG}_${c}_launches_summary.txt 2>&1
echo "== $c"; head -16 gpurun_out/${TAG}_${c}_launches_summary.txt; tail -1 gpurun_out/${TAG}_${c}_launches_summary.txt
done
