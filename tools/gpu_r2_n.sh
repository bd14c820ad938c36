cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/step_timeline.py gpurun_out/r2_timeline.csv 2>&1 | tail -5
head -3 gpurun_out/r2_timeline.csv
