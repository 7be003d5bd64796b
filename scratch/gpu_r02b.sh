#!/bin/bash
mkdir -p gpurun_out
( python scratch/parity_diag.py 64 relu; python scratch/parity_diag.py 8 relu list; python scratch/parity_diag.py 16 tanh list ) > gpurun_out/r02_parity_diag.txt 2>&1
timeout 1500 python -m pytest tests -q -s -m gpu -x > gpurun_out/r02b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest.log
grep -E "^\[|passed|failed|Error|rc=" gpurun_out/r02b_pytest.log | tail -30
