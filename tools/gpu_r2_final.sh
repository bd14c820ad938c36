cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --tb=short ) > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02_gpu_tests.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02_smoke.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 12 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm exit $?"
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench exit $?"
python - <<PY
import json
for f in ("gpurun_out/r02_bench_1gpu.json", "gpurun_out/r02_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("gpu_launches"),
              (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), d.get("loss_check", {}).get("rel"), d.get("clocks"))
    except Exception as e:
        print(f, "parse error", e)
PY
