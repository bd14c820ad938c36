cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -m gpu --tb=short > gpurun_out/r02_pipeline_tests.log 2>&1; echo "pipeline tests exit $?"; tail -5 gpurun_out/r02_pipeline_tests.log | cut -c1-300
lscpu | grep -E "Model name|avx" | cut -c1-200 | head -3; python -c "import cv2; print(cv2.getCPUFeaturesLine())"
for spec in "1 bn" "1 dgrad" "2 bn" "3 bn" "3 dgrad" "4 bn"; do
  set -- $spec
  CILRS_B200_LIB=tools/libcilrs_trace.so CILRS_FLAT_DEBUG=1 TRACE_MAX=1200 timeout 120 python tools/trace_flat.py $1 128 $2 > gpurun_out/trace_l$1_$2.txt 2>&1
  echo "trace $spec exit $?"; grep -m1 "cilrs flat" gpurun_out/trace_l$1_$2.txt | cut -c1-250; grep -m1 "^layer" gpurun_out/trace_l$1_$2.txt
done
