cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in "3 2" "6 2" "3 4" "6 4" "12 8" "2 2"; do
set -- $c
CILRS_EW_FWD_CAP=$1 CILRS_EW_BWD_CAP=$2 timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo -n "fwd_cap=$1 bwd_cap=$2 exit $? "
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['roofline']['breakdown_ms']['bn_forward_pool']['ms'], d['roofline']['breakdown_ms']['bn_backward']['ms'])
PY
done
