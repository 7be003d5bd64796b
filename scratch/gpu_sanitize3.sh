#!/bin/bash
# memcheck over the captured three-stream cGAN step and the cfg3-style tail wiring (padded concatenations)
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider -k "cgan_step or cganstep or CGANStep or cgan" > gpurun_out/r02_sanitize_memcheck_cgan.full.log 2>&1
echo "rc=$?" >> gpurun_out/r02_sanitize_memcheck_cgan.full.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|Program hit" gpurun_out/r02_sanitize_memcheck_cgan.full.log | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_sanitize_memcheck_cgan.log
cat gpurun_out/r02_sanitize_memcheck_cgan.log
