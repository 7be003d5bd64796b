#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/wg_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/wg_pytest.log
tail -8 gpurun_out/wg_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err; cat gpurun_out/r01c_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r01c_ncu_launch.log 2>&1
python scratch/summarize_launches.py gpurun_out/r01c_launches.csv | head -32
