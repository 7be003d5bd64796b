#!/bin/bash
# compute-sanitizer evidence: memcheck and racecheck over the mbarrier / TMEM / TMA kernels and the
# round-2 additions.  Summaries land in gpurun_out/r02_sanitize_*.log (copied to profiles/).
mkdir -p gpurun_out
SEL='test_conv_tc or test_wgrad_stacked_taps or test_subpixel_transition_composed or test_thin_wgrad or test_conv_tc_residual_and_d2s'
run() {   # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider \
      > gpurun_out/r02_sanitize_${tag}.full.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_sanitize_${tag}.full.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Invalid|hazard|Race reported|=========     at" gpurun_out/r02_sanitize_${tag}.full.log | sort | uniq -c | sort -rn | head -40 \
      > gpurun_out/r02_sanitize_${tag}.log
  echo "== $tool $tag"; cat gpurun_out/r02_sanitize_${tag}.log
}
run memcheck memcheck_tc tests/test_gpu_engine.py -k "$SEL"
run racecheck racecheck_tc tests/test_gpu_engine.py -k "(test_conv_tc or test_wgrad_stacked_taps) and (tf32x3 or f16x3)"
run memcheck memcheck_losses tests/test_gpu_losses.py -k "module_functions or spatiotemporal"
run memcheck memcheck_new tests/test_gpu_metrics.py tests/test_gpu_resize.py tests/test_gpu_convnext.py -k "rmse or compute_metrics or resize_op or layer_scale"
