#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_norm.py tests/test_gpu_losses.py -m gpu -q -rf --no-header > gpurun_out/norm_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/norm_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/norm_pytest.log | cut -c1-300 | tail -40
