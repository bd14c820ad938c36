cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -m gpu --tb=short > gpurun_out/r2_tests_l.log 2>&1
tail -5 gpurun_out/r2_tests_l.log | cut -c1-300
timeout 300 python tools/k0_bench.py > gpurun_out/r2_k0.log 2>&1; cat gpurun_out/r2_k0.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
