cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu --tb=short > gpurun_out/r2_pipeline_tests.log 2>&1; tail -25 gpurun_out/r2_pipeline_tests.log | cut -c1-400
CMD="python tools/prof_step.py 3"
LIGHT="ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -f"
FULL="ncu --set full --clock-control none --import-source on -f"
timeout 300 $CMD > gpurun_out/r2_prof_plain.log 2>&1 || { tail -5 gpurun_out/r2_prof_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1; tail -1 gpurun_out/r2_ncu_l.log | cut -c1-200
light() { timeout 900 $LIGHT -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r2_ncu_$1.log 2>&1; tail -1 gpurun_out/r2_ncu_$1.log | cut -c1-160; }
full()  { timeout 900 $FULL  -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r2_ncu_$1.log 2>&1; tail -1 gpurun_out/r2_ncu_$1.log | cut -c1-160; }
light conv_flat_all "conv_flat" 58 58
full  conv_flat_l1 "conv_flat" 62 2
light wgrad_flat_all "wgrad_flat_kernel" 29 29
full  others "wgrad_reduce|adam_kernel|heads_fwd|heads_bwd|bn_relu_maxpool|bn_bwd_reduce_kernel<true>|pack_all" 8 8
timeout 300 $FULL -k regex:preprocess_x4 -c 1 -o gpurun_out/r02_prof_k0_x4 python tools/k0_bench.py > gpurun_out/r2_ncu_k0.log 2>&1; tail -1 gpurun_out/r2_ncu_k0.log | cut -c1-160
rm -f gpurun_out/*.log.tmp
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
