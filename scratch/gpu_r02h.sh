#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -5 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print('step ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_step'])
print('roofline', {k: d['roofline'][k] for k in ('kernel','frac','achieved','step_conv_roofline_frac')})
print('per_function', json.dumps(d['roofline']['per_function'], indent=1))
print('cpu', d['cpu_baseline'])
for k, v in d['configs'].items():
    print(k, v if not isinstance(v, dict) else {kk: v[kk] for kk in ('ms_per_step','value','launches_per_step') if kk in v}, v.get('e2e',{}).get('ms_per_step') if isinstance(v, dict) else '', v.get('roofline',{}).get('frac') if isinstance(v, dict) else '')
PY
