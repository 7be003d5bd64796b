#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_api.py -m gpu -q -rf --no-header > gpurun_out/next_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/next_pytest.log
grep -E "^FAILED|^ERROR|^E  |passed|failed|rc=" gpurun_out/next_pytest.log | cut -c1-300 | tail -40
timeout 600 python scratch/next_configs.py > gpurun_out/next_configs.json 2> gpurun_out/next_configs.err; grep -E "dropout|headline" gpurun_out/next_configs.err
