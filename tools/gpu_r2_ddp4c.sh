cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
env timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-extras > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench exit $?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r02_bench_${N}gpu.err | tail -3 | cut -c1-300
python - <<PY
import json
for f in ('gpurun_out/r02_bench_4gpu.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], 'ms/step %.3f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['config']['allreduce_schedule'], d['config']['grad_comm'], d['config']['cuda_graph'], d.get('infer_c5',{}).get('frames_per_s'), d.get('clocks'))
    except Exception as e: print(f, 'parse error', e)
PY
