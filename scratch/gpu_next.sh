#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_convnext.py -m gpu -q -rf --no-header > gpurun_out/next_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/next_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/next_pytest.log | cut -c1-300 | tail -40
timeout 300 python scratch/bench_next_ops.py > gpurun_out/next_ops.json 2> gpurun_out/next_ops.err; tail -3 gpurun_out/next_ops.err; python -c "
import json; d=json.load(open('gpurun_out/next_ops.json'))
for r in d['rows']: print(r['op'], r['us'], r['GBps'], r['frac_hbm_peak'])"
