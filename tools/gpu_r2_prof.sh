cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras"
timeout 300 python tools/k0_bench.py > gpurun_out/r2_k0.log 2>&1; cat gpurun_out/r2_k0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -c 2 -f -o gpurun_out/r02_prof_k0 python tools/k0_bench.py > gpurun_out/r2_ncu_k0.log 2>&1; tail -1 gpurun_out/r2_ncu_k0.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 400 -c 8 -f -o gpurun_out/r02_prof_conv_flat $CMD > gpurun_out/r2_ncu_cf.log 2>&1; tail -1 gpurun_out/r2_ncu_cf.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_flat|wgrad_reduce|adam_kernel|heads_fwd|heads_bwd|bn_apply_kernel" -s 1200 -c 14 -f -o gpurun_out/r02_prof_others $CMD > gpurun_out/r2_ncu_ot.log 2>&1; tail -1 gpurun_out/r2_ncu_ot.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
