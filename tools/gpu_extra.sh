cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_extra.py all > gpurun_out/extra_bench.json 2> gpurun_out/extra_bench.err; tail -3 gpurun_out/extra_bench.err; cat gpurun_out/extra_bench.json
for k in pre adam heads; do
  case $k in pre) RX="preprocess_kernel";; adam) RX="adam_kernel";; heads) RX="heads_";; esac
  timeout 300 python tools/bench_extra.py $k > gpurun_out/extra_plain_$k.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none -k regex:$RX -s 2 -c 3 -f -o gpurun_out/prof_$k python tools/bench_extra.py $k > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-160
done
ls -la gpurun_out/prof_pre.ncu-rep gpurun_out/prof_adam.ncu-rep gpurun_out/prof_heads.ncu-rep
