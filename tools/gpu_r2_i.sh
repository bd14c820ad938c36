cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2_tests_all.log 2>&1
tail -12 gpurun_out/r2_tests_all.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err; tail -5 gpurun_out/r2_bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_i.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['roofline']['breakdown_ms']), d['loss_first'], d['loss_last'], d['gpu_launches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches_i.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_i.log 2>&1
tail -2 gpurun_out/r2_ncu_i.log | cut -c1-200
