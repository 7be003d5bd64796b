#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02t16}
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -k "thin" > gpurun_out/${TAG}_pytest.log 2>&1
tail -15 gpurun_out/${TAG}_pytest.log
for cfg in "DL4DS_THIN_F16=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --configs cfg4,cfg5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms_per_step', d['ms_per_step'], 'cfg5', d['configs']['cfg5']['ms_per_step'])
for r in d.get('per_function',[])[:0]: print(r)
"
done 2>&1 | tee gpurun_out/${TAG}_ab.log
