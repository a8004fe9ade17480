#!/bin/bash
# Full GPU pass: parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -25 gpurun_out/tests_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1500 python bench.py --files ${FILES:-1000} --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -8 gpurun_out/bench.err
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --files 32 --wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_hybrid|k_huff|k_enc_rate|k_enc_analysis|k_enc_pack' -s 20 -c 5 -o gpurun_out/prof_all \
    python bench.py --files 16 --wave 16 --steps 1 --warmup 1 > gpurun_out/ncu_full_bench.log 2>&1
fi
ls -la gpurun_out
