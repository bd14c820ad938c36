cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --tb=short -x -k "wgrad" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_blocks_b128_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -4
for c in 1 0 1 0; do
CILRS_WGRAD_CLUSTER=$c timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2u_bench_$c.json 2> gpurun_out/r2u_bench.err; echo "bench cluster=$c exit $?"; tail -3 gpurun_out/r2u_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2u_bench_$c.json').read().strip().splitlines()[-1])
print('cluster $c: ms/step', d['ms_per_step'], 'fps', d['value'], 'e2e', d['e2e']['value'], 'wgrad eager ms', d['roofline']['breakdown_ms']['conv_wgrad'])
PY
done
