#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02cfg5f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_cfg5_launches.csv python scratch/prof_cfg5.py > gpurun_out/${TAG}_cfg5_ncu.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_cfg5_launches.csv > gpurun_out/${TAG}_cfg5_launches_summary.txt 2>&1
head -12 gpurun_out/${TAG}_cfg5_launches_summary.txt; tail -1 gpurun_out/${TAG}_cfg5_launches_summary.txt
