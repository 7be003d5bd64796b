#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02f8}
timeout 1500 python -m pytest tests/test_gpu_engine.py tests/test_gpu_baseline_configs.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
for cfg in "DL4DS_NO_FUSED_DGRAD=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --configs cfg3,cfg4,cfg5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms_per_step', d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['configs'].items()})
"
done 2>&1 | tee gpurun_out/${TAG}_ab.log
