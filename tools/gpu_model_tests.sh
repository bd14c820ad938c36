cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/model_test.log
for f in tests/test_bn_heads_gpu.py tests/test_model_gpu.py; do
  echo "=== $f ===" >> gpurun_out/model_test.log
  timeout 600 python -m pytest $f -q -m gpu -s --tb=short 2>&1 | tail -150 >> gpurun_out/model_test.log
done
tail -120 gpurun_out/model_test.log


