#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi.txt
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/multi_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/multi_pytest.log
tail -4 gpurun_out/multi_pytest.log
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err
for f in gpurun_out/scale_n1.json gpurun_out/scale_n2.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"; done
tail -3 gpurun_out/scale_n2.err
