#!/bin/bash
# GPU parity tests only.  usage: tools/gpu_tests.sh [pytest args]
mkdir -p gpurun_out
python -m pytest tests -m gpu -q "$@" > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -60 gpurun_out/tests_gpu.log
