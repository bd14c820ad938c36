cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bn_heads_gpu.py -q -m gpu --tb=short -k "heads" > gpurun_out/r2_tests_heads.log 2>&1
tail -12 gpurun_out/r2_tests_heads.log
timeout 900 python -m pytest tests/test_fp32_mode_gpu.py -q -m gpu -s --tb=short > gpurun_out/r2_tests_fp32.log 2>&1
tail -40 gpurun_out/r2_tests_fp32.log
timeout 900 python -m pytest tests/test_recipe_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short > gpurun_out/r2_tests_recipe.log 2>&1
tail -12 gpurun_out/r2_tests_recipe.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -5 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['roofline']['breakdown_ms']), d['heads'], d['infer_b1'])
PY
