#!/bin/bash
# A/B of environment switches on one box:  AB="VAR=a VAR=b" CONFIGS=cfg5 bash scratch/gpu_ab.sh
mkdir -p gpurun_out
TAG=${TAG:-r02ab}
for rep in 1 2; do
for cfg in $AB; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --configs ${CONFIGS:-cfg5} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms_per_step', d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['configs'].items()})
"
done
done 2>&1 | tee gpurun_out/${TAG}_ab.log
