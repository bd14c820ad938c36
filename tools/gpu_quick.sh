cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --tb=short 2>&1 | tail -8
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench.json"))
print("value %.0f frames/s  %.3f ms/step  e2e %.0f  infer p50 %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j.get("infer_b1")))
print("roofline frac %.3f achieved %.0f TF; step-level sustained frac %.3f" % (j["roofline"]["frac"], j["roofline"]["achieved"], j["roofline"]["step_level"]["frac_sustained"]))
print(j["roofline"]["breakdown_ms"])
PY
