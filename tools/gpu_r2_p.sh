cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bn_heads_gpu.py tests/test_blocks_b128_gpu.py tests/test_model_gpu.py tests/test_recipe_gpu.py -q -m gpu --tb=short -x > gpurun_out/r2p_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r2p_tests.log | cut -c1-400
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2p_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print('bench: ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['breakdown_ms'])
PY
timeout 600 python tools/step_timeline.py gpurun_out/r2p_timeline.csv 2>&1 | tail -1
