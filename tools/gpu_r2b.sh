#!/bin/bash
# round 2, pass b: GPU tests + decode-only bench, fast hybrid kernel vs the direct-form one
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -30 gpurun_out/tests_gpu.log
for mode in fast direct; do
  if [ $mode = direct ]; then export M3S_HYBRID_DIRECT=1; else unset M3S_HYBRID_DIRECT; fi
  timeout 900 python bench.py --steps 3 --warmup 2 --no-encode --no-extras > gpurun_out/bench_dec_$mode.json 2> gpurun_out/bench_dec_$mode.err; echo "bench $mode exit $?"
  python - $mode <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_dec_%s.json" % sys.argv[1]))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"]))
    print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, "frac", d["roofline"]["frac"])
    print(" check", d["check"])
except Exception as ex:
    print("bench parse failed:", ex)
PY
  tail -3 gpurun_out/bench_dec_$mode.err
done
