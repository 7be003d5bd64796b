#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_norm.py tests/test_gpu_convnext.py tests/test_gpu_dropout.py tests/test_gpu_losses.py -m gpu -q -rf --no-header > gpurun_out/next_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/next_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/next_pytest.log | cut -c1-300 | tail -40
timeout 300 python scratch/bench_next_ops.py > gpurun_out/next_ops.json 2> gpurun_out/next_ops.err; tail -3 gpurun_out/next_ops.err; python -c "
import json; d=json.load(open('gpurun_out/next_ops.json'))
for r in d['rows']: print(r['op'], r['us'], r['GBps'], r['frac_hbm_peak'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ssim|range|avgpool" -c 40 --csv --log-file gpurun_out/next_ssim_launches.csv python scratch/bench_next_ops.py > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/next_ssim_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:14]: print(r[4][:40], r[-1])
P
