cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
timeout 900 python -m pytest tests/test_ddp_multigpu.py -q -m gpu -s --tb=short > gpurun_out/r02_ddp_test_$N.log 2>&1
grep -n "passed\|failed\|Error\|error" gpurun_out/r02_ddp_test_$N.log | cut -c1-300 | tail -10
env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-extras > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench exit $?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r02_bench_${N}gpu.err | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_1gpu_samebox.json 2>/dev/null
python - <<PY
import json
for f in ('gpurun_out/r02_bench_2gpu.json','gpurun_out/r02_bench_1gpu_samebox.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], 'ms/step %.3f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['config']['allreduce_schedule'], d['config']['grad_comm'], d['config']['cuda_graph'], d.get('infer_c5',{}).get('frames_per_s'))
    except Exception as e: print(f, 'parse error', e)
PY
