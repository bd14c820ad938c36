cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu --tb=short -x ) > gpurun_out/r2r_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/r2r_tests.log | cut -c1-300
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2r_bench$i.json 2> gpurun_out/r2r_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2r_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench$i.json').read().strip().splitlines()[-1])
print('bench: ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['breakdown_ms'])
PY
done
