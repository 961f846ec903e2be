#!/bin/bash
# developer tool, runs ON the GPU box: round-2 pass P -- axis-aligned kernels, compute-sanitizer, degenerate goldens
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dev_axis.py 2>&1 | tee gpurun_out/r2p_axis.log
timeout 600 python -m pytest tests -m gpu -q -k "degenerate or golden or fast_mode or axis_aligned" > gpurun_out/r2p_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_tests.log; tail -4 gpurun_out/r2p_tests.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "tma_staged or axis_aligned_fp32 or golden_vectors_f64 or (f32_kernel_matches_oracle and 512) or (fast_mode_fp32 and 300) or canvas_taller or separable_tma_path" \
    > gpurun_out/r2p_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -c "ERROR SUMMARY" gpurun_out/r2p_sanitizer_$tool.log; tail -3 gpurun_out/r2p_sanitizer_$tool.log
done
