#!/bin/bash
# round-2 evidence: launch lists (headline bench, cfg5 step), ncu --set full of the headline layers
mkdir -p gpurun_out; TAG=${TAG:-r02w}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 420 --csv --log-file gpurun_out/${TAG}_launches_tf32x3.csv python bench.py --steps 3 --warmup 3 --no-cpu --configs "" > gpurun_out/${TAG}_ncu_launch.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_launches_tf32x3.csv > gpurun_out/${TAG}_launches_tf32x3_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_cfg5_launches.csv python scratch/prof_cfg5.py > gpurun_out/${TAG}_cfg5_ncu.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_cfg5_launches.csv > gpurun_out/${TAG}_cfg5_launches_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|thin_" -c 18 -o gpurun_out/${TAG}_layers -f python scratch/prof_layers.py tf32x3 > gpurun_out/${TAG}_ncu_full.log 2>&1
head -30 gpurun_out/${TAG}_launches_tf32x3_summary.txt; head -30 gpurun_out/${TAG}_cfg5_launches_summary.txt
