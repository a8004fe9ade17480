#!/bin/bash
# GPU pass r1e: parity tests, smoke, pipe microbenchmark, bench, ncu launch list, ncu --set full of decode and encode kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.txt 2>&1
( nproc; free -g | head -2 ) >> gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
./tools/ubench_pipes > gpurun_out/ubench_pipes.txt 2>&1; cat gpurun_out/ubench_pipes.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -4 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --files 32 --wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff|k_walk' -s 6 -c 3 -o gpurun_out/prof_dec \
    python bench.py --files 16 --wave 16 --steps 1 --warmup 1 --no-encode > gpurun_out/ncu_dec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_rate|k_enc_analysis|k_enc_pack' -s 6 -c 3 -o gpurun_out/prof_enc \
    python bench.py --files 1000 --frames 100 --wave 500 --steps 1 --warmup 1 > gpurun_out/ncu_enc.log 2>&1
ls -la gpurun_out
