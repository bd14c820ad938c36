cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
for k in "fprop_matches" "fprop_stats" "stem_fprop" "dgrad" "test_wgrad" "stem_wgrad"; do
  echo "=== $k ===" >> gpurun_out/conv_test.log
  timeout 300 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "$k" -x --tb=short 2>&1 | tail -60 >> gpurun_out/conv_test.log
done
tail -150 gpurun_out/conv_test.log
