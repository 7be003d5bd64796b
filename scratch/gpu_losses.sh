#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_losses.py -m gpu -q -rf --no-header > gpurun_out/losses_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/losses_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/losses_pytest.log | cut -c1-400 | tail -60
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_losses.py -m gpu -q -x -k "module_functions or spatiotemporal" > gpurun_out/losses_sanitizer.log 2>&1; tail -5 gpurun_out/losses_sanitizer.log
