cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 ./tools/heads_trace_test 128 > gpurun_out/r2_heads_trace.log 2>&1; tail -15 gpurun_out/r2_heads_trace.log
timeout 600 python -m pytest tests/test_bn_heads_gpu.py -q -m gpu --tb=short -k "heads" > gpurun_out/r2_tests_heads.log 2>&1
tail -3 gpurun_out/r2_tests_heads.log
timeout 900 python -m pytest tests/test_fp32_mode_gpu.py tests/test_recipe_gpu.py -q -m gpu -s --tb=short > gpurun_out/r2_tests_fp32.log 2>&1
grep -n "fp32 mode\|per-tensor\|passed\|failed\|Error" gpurun_out/r2_tests_fp32.log | cut -c1-400
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -5 gpurun_out/r2_bench_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['roofline']['breakdown_ms']), d['heads'], d['infer_b1'])
PY
