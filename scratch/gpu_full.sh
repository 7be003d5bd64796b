#!/bin/bash
mkdir -p gpurun_out; TAG=${TAG:-r02full}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -12 gpurun_out/${TAG}_pytest_gpu.log
AB="DL4DS_X=0" CONFIGS=cfg3,cfg4,cfg5 bash scratch/gpu_ab.sh
