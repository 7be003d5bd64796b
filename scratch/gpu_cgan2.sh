#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02cg2}
timeout 900 python -m pytest tests -x -q -m gpu -k "cgan or gan or discriminator" > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
for cfg in "DL4DS_CGAN_STREAMS=2" "DL4DS_CGAN_STREAMS=3"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --configs cfg5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms_per_step', d['ms_per_step'], 'cfg5', d['configs']['cfg5']['ms_per_step'], 'e2e', d['configs']['cfg5']['e2e']['ms_per_step'])
"
done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/${TAG}_cfg5_launches.csv python scratch/prof_cfg5.py > gpurun_out/${TAG}_cfg5_ncu.log 2>&1
python scratch/summarize_launches.py gpurun_out/${TAG}_cfg5_launches.csv > gpurun_out/${TAG}_cfg5_launches_summary.txt 2>&1
