cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short -x -k "wgrad" 2>&1 | tail -3
for c in 148 111 96 74 48; do
CILRS_WGRAD_CTAS=$c timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_$c.json 2> gpurun_out/r2w_bench.err; echo "bench ctas=$c exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2w_bench_$c.json').read().strip().splitlines()[-1])
print('ctas $c: ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], 'wgrad eager ms', d['roofline']['breakdown_ms']['conv_wgrad'])
PY
done
