cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "3 wgrad wgrad_flat" "3 bn conv_flat" "2 dgrad conv_flat"; do
  set -- $cfg
  timeout 120 python tools/prof_flat.py $1 $2 > gpurun_out/pf_plain.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o gpurun_out/prof_flat_l$1_$2 python tools/prof_flat.py $1 $2 > gpurun_out/ncu_flat_l$1_$2.log 2>&1
  tail -2 gpurun_out/ncu_flat_l$1_$2.log
done
timeout 120 python tools/bench_ew.py > gpurun_out/ew_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:bn_apply_kernel -s 2 -c 2 -f -o gpurun_out/prof_bn_apply python tools/bench_ew.py > gpurun_out/ncu_bn_apply.log 2>&1
tail -2 gpurun_out/ncu_bn_apply.log
ls -la gpurun_out/*.ncu-rep
