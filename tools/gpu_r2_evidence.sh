# Round-2 evidence run (one GPU): bench (both arms), the GPU test suite, launch list, ncu captures of the hot kernels, step timeline.
# Everything lands under gpurun_out/; tools/collect_profiles.py turns it into profiles/r02_*.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
# ---- 1. bench: our arm, then the reference arm (same config) ----
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; el "bench exit $?"
timeout 600 python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; el "reference arm exit $?"
python - <<PY
import json
for f in ("gpurun_out/r02_bench_1gpu.json", "gpurun_out/r02_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("gpu_launches"),
              (d.get("roofline") or {}).get("frac"), d.get("loss_check", {}).get("rel"), d.get("clocks"))
    except Exception as e:
        print(f, "parse error", e)
PY
# ---- 2. the GPU test suite, as the driver runs it ----
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --tb=short ) > gpurun_out/r02_gpu_tests.log 2>&1; el "pytest exit $?"; tail -6 gpurun_out/r02_gpu_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; el "smoke exit $?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-300
# ---- 3. step timeline (CUPTI) ----
timeout 300 python tools/step_timeline.py gpurun_out/r02_step_timeline.csv 2>&1 | tail -1
# ---- 4. ncu: launch list of three eager steps, then section captures ----
CMD="python tools/prof_step.py 3"
timeout 300 $CMD > gpurun_out/r2_prof_plain.log 2>&1 || { tail -5 gpurun_out/r2_prof_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1; el "launch list exit $?"
LIGHT="ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -f"
FULL="ncu --set full --clock-control none --import-source on -f"
light() { timeout 600 $LIGHT -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r2_ncu_$1.log 2>&1; el "$1 exit $?"; }
full()  { timeout 600 $FULL  -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 $CMD > gpurun_out/r2_ncu_$1.log 2>&1; el "$1 exit $?"; }
# (skip the first step's launches: the third step is profiled)
light conv_flat_all "conv_flat" 116 58
full  conv_flat_l1 "conv_flat" 120 2
light wgrad_flat_all "wgrad_flat_kernel" 58 29
FULL="ncu --set full --clock-control none -f"
full  others "wgrad_reduce|adam_kernel|heads_fwd|heads_bwd|heads_wgrad|bn_relu_maxpool|bn_apply_kernel|bn_bwd_apply|conv_gemm_kernel|wgrad_gemm" 40 20
timeout 300 $FULL -k regex:preprocess_x4 -c 1 -o gpurun_out/r02_prof_k0_x4 python tools/k0_bench.py > gpurun_out/r2_ncu_k0.log 2>&1; el "k0 exit $?"
timeout 120 python tools/k0_bench.py > gpurun_out/r02_k0_bench.json 2>&1; cat gpurun_out/r02_k0_bench.json | cut -c1-300
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
el done
