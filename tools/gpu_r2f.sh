#!/bin/bash
# round 2, pass f -- the full single-GPU pass behind profiles/r2f_*: GPU tests, smoke, bench (both arms), ncu launch list,
# ncu --set full of the decode and encode kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/tests_gpu.log; tail -3 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cut -c1-300 gpurun_out/bench_reference.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"], d["roofline"]["frac"]))
    print(" kernels", {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    e = d["encode_hide"]
    print("encode value %.4g e2e %.4g ms %.1f e2e_ms %.1f" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["e2e"]["ms_per_step"]))
    print(" check", d["check"]["parity_sampled"], "composite", d.get("composite"))
    c = d["cfg5"]; print(" cfg5 value %.4g e2e %.4g parity %s long %s" % (c["value"], c["e2e"]["value"], c["parity_ok"], c["long_file"]))
    print(" cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("python_reference"))
except Exception as ex:
    print("bench parse failed:", ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --files 64 --steps 1 --warmup 1 --no-extras > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff|k_walk|k_strip' -s 8 -c 4 -f -o gpurun_out/prof_dec \
    python bench.py --files 32 --steps 1 --warmup 1 --no-encode --no-extras > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec exit $?"
python tools/ncu_pick.py gpurun_out/prof_dec.ncu-rep gpurun_out/dec_ncu_full_summary.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_' -s 10 -c 5 -f -o gpurun_out/prof_enc \
    python bench.py --files 1000 --frames 100 --steps 1 --warmup 1 --no-extras > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc exit $?"
python tools/ncu_pick.py gpurun_out/prof_enc.ncu-rep gpurun_out/enc_ncu_full_summary.csv
ls -la gpurun_out | head -40
