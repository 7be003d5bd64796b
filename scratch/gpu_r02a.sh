#!/bin/bash
# round 2, call a: UMMA rate re-probe + the literal-config parity tests + the resume test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 120 scratch/umma_rate2 > gpurun_out/r02a_umma_rate2.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_umma_rate2.log
timeout 60 scratch/umma_rate2 cg2 > gpurun_out/r02a_umma_rate2_cg2.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_umma_rate2_cg2.log
tail -3 gpurun_out/r02a_umma_rate2.log gpurun_out/r02a_umma_rate2_cg2.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_api.py -q -s -m gpu > gpurun_out/r02a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_pytest.log
grep -E "^\[|passed|failed|Error|error|rc=" gpurun_out/r02a_pytest.log | tail -40
