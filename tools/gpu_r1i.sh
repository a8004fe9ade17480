#!/bin/bash
# GPU pass r1i: all GPU tests, bench (both arms), ncu launch list, ncu --set full of the decode and encode kernels (parallel rate loop)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/tests_gpu.log; tail -3 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cat gpurun_out/bench_reference.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --files 64 --wave 32 --e2e-wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff|k_walk' -s 6 -c 3 -f -o gpurun_out/prof_dec \
    python bench.py --files 16 --wave 16 --steps 1 --warmup 1 --no-encode > gpurun_out/ncu_dec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_' -s 10 -c 5 -f -o gpurun_out/prof_enc \
    python bench.py --files 1000 --frames 100 --wave 500 --steps 1 --warmup 1 > gpurun_out/ncu_enc.log 2>&1
ls -la gpurun_out
