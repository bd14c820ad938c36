cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_blocks_b128_gpu.py tests/test_recipe_gpu.py -q -m gpu --tb=short -x > gpurun_out/r2_tests_k.log 2>&1
tail -8 gpurun_out/r2_tests_k.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err; tail -5 gpurun_out/r2_bench_k.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_k.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['roofline']['breakdown_ms']), d['loss_first'], d['loss_last'], d['gpu_launches'])
PY
