cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 900 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -3
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -1 gpurun_out/ncu1.log | cut -c1-200
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 400 -c 8 -f -o gpurun_out/prof_conv_flat $CMD > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200
timeout 600 $CMD > gpurun_out/plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:"wgrad_flat|bn_bwd_apply|bn_apply_kernel|adam_kernel|preprocess_kernel|heads_bwd" -s 1200 -c 12 -f -o gpurun_out/prof_others $CMD > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/ncu3.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
