cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python tools/prof_step.py 3"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"bn_relu_maxpool_sel|stem_bwd_apply|pack_all|adam_kernel|normalize_s2d" -s 10 -c 5 -o gpurun_out/r02_prof_stem_new $CMD > gpurun_out/r2_ncu_stem_new.log 2>&1; echo "stem_new exit $?"
ls -la gpurun_out/r02_prof_stem_new*
