#!/bin/bash
mkdir -p gpurun_out; TAG=${TAG:-r02hl}
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 420 --csv --log-file gpurun_out/${TAG}_launches_tf32x3.csv python bench.py --steps 3 --warmup 3 --no-cpu --configs "" > gpurun_out/${TAG}_ncu_launch.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_launches_tf32x3.csv > gpurun_out/${TAG}_launches_tf32x3_summary.txt
head -32 gpurun_out/${TAG}_launches_tf32x3_summary.txt; tail -1 gpurun_out/${TAG}_launches_tf32x3_summary.txt
