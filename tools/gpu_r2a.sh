#!/bin/bash
# round 2, pass a: GPU tests + smoke + bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -25 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"]))
    print(" host roof", d["e2e"]["host_roof_gbs"])
    print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    e = d.get("encode_hide")
    if e:
        print("encode value %.4g e2e %.4g ms %.1f e2e_ms %.1f" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["e2e"]["ms_per_step"]))
        print(" kernels", {k: round(v, 1) for k, v in e["roofline"]["kernel_ms_per_step"].items()})
    print(" check", d["check"]); print(" composite", d.get("composite")); print(" cfg5", d.get("cfg5"))
    print(" cpu", d["cpu_baseline"])
except Exception as ex:
    print("bench parse failed:", ex)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"; cut -c1-600 gpurun_out/bench_reference.json
