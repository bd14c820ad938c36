cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; tail -1 gpurun_out/ddp_check_$N.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_$N.json 2>gpurun_out/scale_$N.err; echo "exit $?"; tail -2 gpurun_out/scale_$N.err | cut -c1-300; cat gpurun_out/scale_$N.json | cut -c1-700
