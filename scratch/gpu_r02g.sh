#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r02g_n1.json 2> gpurun_out/r02g_n1.err; tail -2 gpurun_out/r02g_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r02g_n2.json 2> gpurun_out/r02g_n2.err; tail -3 gpurun_out/r02g_n2.err
python -c "
import json
for n in (1,2):
    d=json.load(open('gpurun_out/r02g_n%d.json'%n)); print(n, 'step ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
