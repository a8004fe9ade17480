#!/bin/bash
# First GPU pass of a round: parity tests, a bench line, the ncu launch list and one full capture of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/box.txt 2>&1
(nproc; free -g | head -2) >> gpurun_out/box.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --files ${FILES:-1000} --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --files 32 --wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff|k_strip' -s 3 -c 3 -o gpurun_out/prof_decode \
    python bench.py --files 16 --wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
