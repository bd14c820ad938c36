cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in 3 4; do
CILRS_BN_FUSION=$mode timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_fuse$mode.json 2> gpurun_out/r2_bench_fuse$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_fuse$mode.json').read().strip().splitlines()[-1])
print('fusion mode $mode:', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], json.dumps(d['roofline']['breakdown_ms']))
PY
done
