#!/bin/bash
# r2k: GPU tests, smoke and the bench line with the folded analysis kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/tests_gpu.log; tail -3 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"], d["roofline"]["frac"]))
print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
e = d["encode_hide"]
print("encode value %.4g e2e %.4g ms %.1f e2e_ms %.1f" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["e2e"]["ms_per_step"]), e.get("roofline"))
print(" check", d["check"], "composite", d.get("composite"))
PY
