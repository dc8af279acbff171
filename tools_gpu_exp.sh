#!/bin/bash
mkdir -p gpurun_out
# compute-sanitizer memcheck over a small, representative subset of the GPU tests (every kernel family once)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests -m gpu -q --timeout 600 -x -k "test_update_kernel_matches_oracle and 256 or test_and_update_matches_oracle and 256 or test_philox or test_n4 and a or test_attention_core_matches_oracle and 256 or test_k3_self or test_conv3x3_fused_2cta_kernel and 16-8 or test_unet_forward_matches_oracle_and_golden and 16" > gpurun_out/sanitizer_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_pytest.log; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/sanitizer_memcheck.log | head -10
