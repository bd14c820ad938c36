cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 ./tools/heads_trace_test 128 > gpurun_out/r2_heads_trace.log 2>&1; cat gpurun_out/r2_heads_trace.log
timeout 120 ./tools/heads_trace_test 1 >> gpurun_out/r2_heads_trace.log 2>&1; tail -16 gpurun_out/r2_heads_trace.log
timeout 900 python -m pytest tests/test_fp32_mode_gpu.py -q -m gpu -s --tb=short > gpurun_out/r2_tests_fp32.log 2>&1
grep -n "fp32 mode\|per-tensor\|passed\|failed" gpurun_out/r2_tests_fp32.log | cut -c1-500
