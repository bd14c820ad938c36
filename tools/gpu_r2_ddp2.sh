cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_ddp_multigpu.py -q -m gpu -s --tb=short > gpurun_out/r2_ddp_test_$N.log 2>&1
grep -n "ddp world\|passed\|failed\|Error\|error" gpurun_out/r2_ddp_test_$N.log | cut -c1-300 | tail -40
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-extras > gpurun_out/r2_scale_${N}_$tag.json 2> gpurun_out/r2_scale_${N}_$tag.err
  echo "$tag exit $?"; tail -2 gpurun_out/r2_scale_${N}_$tag.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', d['n_gpus'], 'ms/step %.3f'%d['ms_per_step'], 'fps %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['config']['allreduce_schedule'], d['config']['grad_comm'], d['config']['cuda_graph'])
except Exception as e: print('$tag parse error', e)
PY
}
run two_bf16 CILRS_BENCH_ALLREDUCE=two CILRS_BENCH_GRAD_COMM=bf16
run first_fp32 CILRS_BENCH_ALLREDUCE=first CILRS_BENCH_GRAD_COMM=fp32
run all_bf16 CILRS_BENCH_ALLREDUCE=all CILRS_BENCH_GRAD_COMM=bf16
run two_bf16_async CILRS_BENCH_ALLREDUCE=two CILRS_BENCH_GRAD_COMM=bf16 CILRS_BENCH_ASYNC_PARTS=1
