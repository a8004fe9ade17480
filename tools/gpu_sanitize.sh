#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on everything, racecheck on the shared-memory heavy kernels' tests,
# initcheck (uninitialised device reads: the main-data stream S is no longer cleared) on the decode tests
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --deselect tests/test_integration_stub.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py "tests/test_parity_encode.py::test_batch_vs_oracle" \
    "tests/test_parity_encode.py::test_quiet_silent_and_loud" "tests/test_parity_encode.py::test_chunked_equals_single" -m gpu -x -q > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?"; tail -3 gpurun_out/racecheck.log
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/initcheck.log 2>&1
echo "initcheck exit $?"; tail -3 gpurun_out/initcheck.log; grep -c "Uninitialized" gpurun_out/initcheck.log
