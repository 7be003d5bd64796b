#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -k "thin" 2>&1 | tail -3
AB="DL4DS_X=0" bash scratch/gpu_ab.sh
TAG=r02cfg5h bash scratch/gpu_cfg5b.sh > /dev/null
grep -i "thin_wgrad_kernel\|thin_conv_kernel" gpurun_out/r02cfg5h_cfg5_launches_summary.txt
