#!/bin/bash
# compute-sanitizer over the small parity cases of the round-2 kernels (limb path, BRDFs, multi-SZA BVP, TMA row staging,
# validation): memcheck for out-of-bounds / misaligned accesses, racecheck for shared-memory hazards.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_limb.py tests/test_brdf.py tests/test_validation.py -m gpu -x -q \
      -k "not chunked" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3 | tee -a gpurun_out/sanitize_summary.txt
done
SK_B200_BVP_TMA=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_counts or config1_shape" > gpurun_out/sanitize_tma.log 2>&1
echo "tma memcheck rc=$?" | tee -a gpurun_out/sanitize_summary.txt
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_tma.log | tail -2 | tee -a gpurun_out/sanitize_summary.txt
