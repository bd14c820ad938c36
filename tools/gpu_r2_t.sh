cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_recipe_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3
for i in 1 2 3; do
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench_$i.json 2> gpurun_out/r2t_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench_$i.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
timeout 600 python tools/step_timeline.py gpurun_out/r2t_timeline.csv 2>&1 | tail -1
