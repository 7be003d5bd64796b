#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:thin_wgrad_kernel -c 4 -o gpurun_out/r02_thinw -f python scratch/prof_cfg5.py > gpurun_out/r02_thinw_ncu.log 2>&1
tail -3 gpurun_out/r02_thinw_ncu.log
ls -la gpurun_out/r02_thinw.ncu-rep
