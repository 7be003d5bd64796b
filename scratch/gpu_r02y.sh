#!/bin/bash
# round-2 closing evidence on one B200: full GPU suite, smoke, both bench arms, fast mode, f16x3 mode
mkdir -p gpurun_out; export TAG=${TAG:-r02y}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_cpu.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_tf32x3.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 30 --warmup 5 --math tf32 --no-cpu --configs "" > gpurun_out/${TAG}_bench_tf32.json 2>/dev/null
timeout 300 python bench.py --steps 30 --warmup 5 --math f16x3 --no-cpu --configs "" > gpurun_out/${TAG}_bench_f16x3.json 2>/dev/null
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log
python - <<'PY'
import json, os
for f in [os.environ.get('TAG', 'r02y') + s for s in ('_bench_tf32x3', '_bench_tf32', '_bench_f16x3', '_bench_reference_cpu')]:
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'),
              {k: v.get('ms_per_step') for k, v in (d.get('configs') or {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
