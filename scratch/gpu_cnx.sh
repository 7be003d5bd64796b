#!/bin/bash
# convnext round: thin 7x7 tests, step time, ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_convnext.py -x -q -k "thin or convnext or wgrad_stacked" > gpurun_out/r02cnx_pytest.log 2>&1
tail -5 gpurun_out/r02cnx_pytest.log
timeout 300 python scratch/prof_convnext.py 2>&1 | tail -3 | tee gpurun_out/r02cnx_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02cnx_convnext_launches.csv python scratch/prof_convnext.py > gpurun_out/r02cnx_ncu.log 2>&1
tail -2 gpurun_out/r02cnx_ncu.log
