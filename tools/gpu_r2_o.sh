cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_recipe_gpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1])
print('bench:', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d.get('loss_check'))
PY
timeout 600 python tools/step_timeline.py gpurun_out/r2_timeline3.csv 2>&1 | tail -1
