#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "wgrad_stacked or conv_tc or spc_block or net_resnet_spc_tc" > gpurun_out/wg_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/wg_pytest.log
tail -15 gpurun_out/wg_pytest.log
timeout 200 python scratch/time_layers.py tf32x3 > gpurun_out/wg_time_new.log 2>&1; echo "rc=$?" >> gpurun_out/wg_time_new.log
cat gpurun_out/wg_time_new.log
timeout 100 python scratch/wg2_stamps.py 64 32 32 48 48 3 > gpurun_out/st_bb48.log 2>&1; timeout 100 python scratch/wg2_stamps.py 64 128 128 48 8 1 > gpurun_out/st_tl.log 2>&1; timeout 100 python scratch/wg2_stamps.py 64 64 64 48 192 3 > gpurun_out/st_spc.log 2>&1; head -20 gpurun_out/st_bb48.log
