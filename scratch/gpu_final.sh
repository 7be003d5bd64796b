#!/bin/bash
mkdir -p gpurun_out; TAG=${TAG:-r02fin}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log


timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_tf32x3.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_tf32x3.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['configs'].items()})"
