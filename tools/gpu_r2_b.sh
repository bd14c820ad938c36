cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_blocks_b128_gpu.py tests/test_recipe_gpu.py tests/test_bn_heads_gpu.py -q -m gpu -s --tb=short > gpurun_out/r2_tests_new.log 2>&1
tail -40 gpurun_out/r2_tests_new.log
timeout 300 python tools/adam_bench.py > gpurun_out/r2_adam_bench.log 2>&1; cat gpurun_out/r2_adam_bench.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -5 gpurun_out/r2_bench_b.err; cat gpurun_out/r2_bench_b.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches_b.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_b.log 2>&1
tail -3 gpurun_out/r2_ncu_b.log
