cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in "4 4 2" "2 4 2" "4 2 2" "4 4 1" "2 2 1"; do
set -- $c
CILRS_EW_POOL_CAP=$1 CILRS_EW_STEM_CAP=$2 CILRS_EW_REDUCE_CAP=$3 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo -n "pool=$1 stem=$2 reduce=$3 exit $? "
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'])
PY
done
