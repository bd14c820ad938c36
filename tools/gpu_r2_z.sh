cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu --tb=short -x ) > gpurun_out/r2z_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/r2z_tests.log | cut -c1-300
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['roofline']['breakdown_ms'])
PY
timeout 600 python tools/step_timeline.py gpurun_out/r2z_timeline.csv 2>&1 | tail -1
