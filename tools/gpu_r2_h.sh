cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_b128_gpu.py -q -m gpu -s --tb=short -x > gpurun_out/r2_tests_fuse.log 2>&1
grep -n "fused vs\|passed\|failed\|Error\|error" gpurun_out/r2_tests_fuse.log | cut -c1-300 | tail -20
grep -n "vs emu" gpurun_out/r2_tests_fuse.log | cut -c1-260 | head -20
timeout 1200 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_blocks_b128_gpu.py > gpurun_out/r2_tests_rest.log 2>&1
tail -15 gpurun_out/r2_tests_rest.log | cut -c1-300
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; tail -5 gpurun_out/r2_bench_h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_h.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], json.dumps(d['roofline']['breakdown_ms']), d['loss_first'], d['loss_last'], d['gpu_launches'])
PY
CILRS_NO_FUSE=1 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_h_nofuse.json 2> gpurun_out/r2_bench_h_nofuse.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_h_nofuse.json').read().strip().splitlines()[-1])
print('nofuse', d['ms_per_step'], d['value'], d['loss_first'], d['loss_last'], d['gpu_launches'])
PY
