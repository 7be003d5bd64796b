#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L > gpurun_out/multi_smi_n${N}.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/scale_n${N}.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d.get('configs',{}).items()})"
tail -2 gpurun_out/scale_n${N}.err
