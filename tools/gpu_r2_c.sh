cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bn_heads_gpu.py -q -m gpu -s --tb=short -k "heads" > gpurun_out/r2_tests_heads.log 2>&1
tail -15 gpurun_out/r2_tests_heads.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_bn_heads_gpu.py -q -m gpu -k "heads_forward_backward" > gpurun_out/r2_sanitizer_heads.log 2>&1
tail -8 gpurun_out/r2_sanitizer_heads.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2_tests_all.log 2>&1
tail -40 gpurun_out/r2_tests_all.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -5 gpurun_out/r2_bench_c.err; cat gpurun_out/r2_bench_c.json
