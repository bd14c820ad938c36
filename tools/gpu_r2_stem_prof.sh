cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python tools/prof_step.py 3"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"bn_relu_maxpool|bn_bwd_reduce_kernel<1>|bn_bwd_reduce_kernel<true>|pack_all|adam_kernel|normalize_s2d" -s 8 -c 10 -o gpurun_out/r02_prof_stem_ew $CMD > gpurun_out/r2_ncu_stem_ew.log 2>&1; echo "stem_ew exit $?"
# the stem's convolution (first conv_gemm launch of a step) and its weight gradient (last wgrad_gemm launch of a step): third step
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"^cilrs::conv_gemm_kernel|^conv_gemm_kernel" -s 14 -c 1 -o gpurun_out/r02_prof_stem_conv $CMD > gpurun_out/r2_ncu_stem_conv.log 2>&1; echo "stem_conv exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"wgrad_gemm_kernel" -s 23 -c 1 -o gpurun_out/r02_prof_stem_wgrad $CMD > gpurun_out/r2_ncu_stem_wgrad.log 2>&1; echo "stem_wgrad exit $?"
# the stem's bn_bwd_apply is the last bn_bwd_apply launch of a step (37 per step)
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"bn_bwd_apply" -s 110 -c 1 -o gpurun_out/r02_prof_stem_bnapply $CMD > gpurun_out/r2_ncu_stem_bnapply.log 2>&1; echo "stem_bnapply exit $?"
ls -la gpurun_out/r02_prof_stem*
