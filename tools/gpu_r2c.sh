#!/bin/bash
# round 2, pass c: GPU tests, ncu launch list + --set full of the decode kernels (fast hybrid), e2e pipeline timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff' -s 4 -c 2 -f -o gpurun_out/prof_dec \
    python bench.py --files 32 --steps 1 --warmup 1 --no-encode --no-extras > gpurun_out/ncu_dec.log 2>&1; echo "ncu exit $?"
python tools/ncu_pick.py gpurun_out/prof_dec.ncu-rep gpurun_out/dec_ncu_full_summary.csv; cat gpurun_out/dec_ncu_full_summary.csv
M3S_TRACE=1 timeout 600 python bench.py --files 300 --steps 1 --warmup 1 --no-encode --no-extras > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo "trace exit $?"
grep -n "m3s_decode\] " gpurun_out/bench_trace.err | tail -2
