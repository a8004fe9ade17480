#!/bin/bash
# quick GPU pass: parity tests + smoke + one bench run.  usage: tools/gpu_quick.sh [bench args]
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -8 gpurun_out/tests_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
    print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    e = d.get("encode_hide")
    if e:
        print("encode value %.4g e2e %.4g ms %.1f e2e_ms %.1f" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["e2e"]["ms_per_step"]))
        print(" kernels", {k: round(v, 1) for k, v in e["roofline"]["kernel_ms_per_step"].items()})
    print(" clocks", d["clocks"])
except Exception as ex:
    print("bench parse failed:", ex)
PY
tail -5 gpurun_out/bench.err
