#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: fp16 warp-level 8-channel kernels (incl. the fused
# epilogue-backward), thin 7x7 / 2- / 4-channel kernels with the transposed final reduction, pointwise kernels with channel
# tails, padded concatenations.  Summaries: gpurun_out/r02_sanitize_*.log (copied to profiles/).
mkdir -p gpurun_out
run() {   # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider \
      > gpurun_out/r02_sanitize_${tag}.full.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitize_${tag}.full.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Invalid|hazard|Race reported|=========     at" gpurun_out/r02_sanitize_${tag}.full.log | sort | uniq -c | sort -rn | head -40 \
      > gpurun_out/r02_sanitize_${tag}.log
  echo "== $tool $tag"; cat gpurun_out/r02_sanitize_${tag}.log
}
run memcheck memcheck_thin16 tests/test_gpu_engine.py -k "test_thin_mma_8x8 or test_thin_wgrad or test_pointwise_conv"
run racecheck racecheck_thin16 tests/test_gpu_engine.py -k "(test_thin_mma_8x8 and tf32x3 and hw1) or (test_thin_wgrad and hw1) or test_pointwise_conv_channel_tails"
