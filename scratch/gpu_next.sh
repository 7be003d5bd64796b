#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_resize.py -m gpu -q -rf --no-header > gpurun_out/next_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/next_pytest.log
grep -E "^FAILED|^ERROR|^E  |passed|failed|rc=" gpurun_out/next_pytest.log | cut -c1-300 | tail -40
