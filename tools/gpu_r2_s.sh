cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pr in 1 0 1 0; do
CILRS_CAPTURE_PRIORITY=$pr timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_$pr.json 2> gpurun_out/r2s_bench.err; echo "bench prio=$pr exit $?"; tail -3 gpurun_out/r2s_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2s_bench_$pr.json').read().strip().splitlines()[-1])
print('prio $pr: ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
done
CILRS_CAPTURE_PRIORITY=1 timeout 600 python tools/step_timeline.py gpurun_out/r2s_timeline.csv 2>&1 | tail -1
timeout 600 python -m pytest tests/test_recipe_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -3
