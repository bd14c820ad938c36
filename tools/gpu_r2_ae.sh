cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_b128_gpu.py tests/test_model_gpu.py tests/test_recipe_gpu.py tests/test_conv_gpu.py -q -m gpu --tb=short -x 2>&1 | tail -5
for c in 0 1 0 1; do
if [ $c = 1 ]; then export CILRS_NO_MASKSUM=1; else unset CILRS_NO_MASKSUM; fi
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo -n "no_masksum=$c exit $? "
python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'loss_last', d['loss_last'], d['roofline']['breakdown_ms']['conv_dgrad'], d['roofline']['breakdown_ms']['bn_backward'])
PY
done
