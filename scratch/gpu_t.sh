#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/q_pytest.log
tail -12 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print('step ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step']); print(d['top_kernels_ms_per_step'])"
DL4DS_TC_NO_T=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_T step ms', d['ms_per_step'])"
