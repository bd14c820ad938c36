cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 148 74 48 111 74 148; do
CILRS_WGRAD_GEMM_CTAS=$c timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo -n "gemm ctas=$c exit $? "
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'wgrad eager ms', d['roofline']['breakdown_ms']['conv_wgrad']['ms'])
PY
done
