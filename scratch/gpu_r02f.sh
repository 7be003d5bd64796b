#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 420 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02f_ncu_launch.log 2>&1
python scratch/summarize_launches.py gpurun_out/r02f_launches.csv | head -40
