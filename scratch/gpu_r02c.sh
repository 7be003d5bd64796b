#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -q -x -m gpu > gpurun_out/r02c_pytest_$name.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_$name.log
  echo "== $name"; grep -E "passed|failed|Error|rc=" gpurun_out/r02c_pytest_$name.log | tail -5
}
run default A=1
if ! grep -q "rc=0" gpurun_out/r02c_pytest_default.log; then
  run baseoff DL4DS_HALO_BASEOFF=1
  run pitch8 DL4DS_HALO_PITCH_ALIGN=8
  run pitch8_baseoff DL4DS_HALO_PITCH_ALIGN=8 DL4DS_HALO_BASEOFF=1
fi
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -2 gpurun_out/r02c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print('step ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step']); print(d['top_kernels_ms_per_step'])"
DL4DS_TC_NO_HALO=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r02c_bench_nohalo.json 2> gpurun_out/r02c_bench_nohalo.err
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_nohalo.json')); print('NO HALO step ms', d['ms_per_step'], 'value', d['value']); print(d['top_kernels_ms_per_step'])"
