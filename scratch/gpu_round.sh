#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu --set full of the headline layers.
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --math tf32 --no-cpu > gpurun_out/${TAG}_bench_tf32.json 2> gpurun_out/${TAG}_bench_tf32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|thin_" -c 18 -o gpurun_out/${TAG}_layers -f python scratch/prof_layers.py tf32x3 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json
timeout 300 python scratch/bench_next_ops.py > gpurun_out/${TAG}_next_ops.json 2> gpurun_out/${TAG}_next_ops.err
timeout 600 python scratch/next_configs.py > gpurun_out/${TAG}_next_configs.json 2> gpurun_out/${TAG}_next_configs.err; tail -3 gpurun_out/${TAG}_next_configs.err
