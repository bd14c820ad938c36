cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  N=$1; tag=$2; shift; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --no-extras --no-cpu-baseline > gpurun_out/r2b_scale_${N}_$tag.json 2> gpurun_out/r2b_scale_${N}_$tag.err
  echo "$tag exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2b_scale_${N}_$tag.err | tail -2 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2b_scale_${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', d['n_gpus'], 'ms/step %.3f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['config']['allreduce_schedule'], d['config']['grad_comm'], d['config'].get('async_parts'), d['config'].get('nccl_max_ctas'), 'c5 %.0f'%d['infer_c5']['frames_per_s'])
except Exception as e: print('$tag parse error', e)
PY
}
run 8 default
run 8 ctas8 CILRS_BENCH_NCCL_MAX_CTAS=8
run 8 async CILRS_BENCH_ASYNC_PARTS=1
run 8 async_ctas8 CILRS_BENCH_ASYNC_PARTS=1 CILRS_BENCH_NCCL_MAX_CTAS=8
