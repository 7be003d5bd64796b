#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python scratch/time_layers.py tf32x3 2>&1 | grep -v Warn | tee gpurun_out/r02e_time_layers.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print('step ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value']); print(d['top_kernels_ms_per_step'])"
