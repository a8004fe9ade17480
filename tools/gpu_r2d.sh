#!/bin/bash
# round 2, pass d: GPU tests, decode-only bench, ncu --set full of the decode kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-encode --no-extras > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_dec.json"))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"]))
    print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, "frac", d["roofline"]["frac"])
    print(" check", d["check"])
except Exception as ex:
    print("bench parse failed:", ex)
PY
tail -3 gpurun_out/bench_dec.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff' -s 4 -c 2 -f -o gpurun_out/prof_dec \
    python bench.py --files 32 --steps 1 --warmup 1 --no-encode --no-extras > gpurun_out/ncu_dec.log 2>&1; echo "ncu exit $?"
python tools/ncu_pick.py gpurun_out/prof_dec.ncu-rep gpurun_out/dec_ncu_full_summary.csv; head -16 gpurun_out/dec_ncu_full_summary.csv; grep "local_op\|thread_inst\|long_score\|barrier" gpurun_out/dec_ncu_full_summary.csv
