#!/bin/bash
# Quick GPU pass: parity tests + one bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 900 python bench.py --files ${FILES:-1000} --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
