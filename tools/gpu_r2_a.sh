cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_blocks_b128_gpu.py tests/test_recipe_gpu.py -q -m gpu -s --tb=short > gpurun_out/r2_tests_new.log 2>&1
tail -60 gpurun_out/r2_tests_new.log
timeout 900 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_blocks_b128_gpu.py --deselect tests/test_recipe_gpu.py > gpurun_out/r2_tests_old.log 2>&1
tail -15 gpurun_out/r2_tests_old.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -5 gpurun_out/r2_bench_a.err; cat gpurun_out/r2_bench_a.json
