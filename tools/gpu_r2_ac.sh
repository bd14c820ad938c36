cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 1 0 1; do
CILRS_BENCH_ADAM_BESIDE_STEM=$c timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo -n "adam_beside_stem=$c exit $? "
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d.get('loss_last'))
PY
done
