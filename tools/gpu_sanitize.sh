#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on everything (incl. the corrupted-input test), racecheck on the shared-memory
# heavy kernels' tests, initcheck (uninitialised device reads: the main-data stream S is not cleared) on the decode tests.
# Summary -> gpurun_out/sanitizer_summary.txt (profiles/r2g_sanitizer.txt)
mkdir -p gpurun_out
S=gpurun_out/sanitizer_summary.txt
echo "compute-sanitizer (CUDA 12.9) over the GPU parity tests on a B200 (tools/gpu_sanitize.sh)" > $S
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --deselect tests/test_integration_stub.py > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $? (all GPU tests but the reference-stub subprocess test)" | tee -a $S; tail -3 gpurun_out/memcheck.log | tee -a $S
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py "tests/test_parity_encode.py::test_batch_vs_oracle" \
    "tests/test_parity_encode.py::test_quiet_silent_and_loud" "tests/test_parity_encode.py::test_chunked_equals_single" -m gpu -x -q > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $? (decode, shard and three encode tests)" | tee -a $S; tail -3 gpurun_out/racecheck.log | tee -a $S
timeout 1500 compute-sanitizer --tool initcheck --print-limit 2000 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/initcheck.log 2>&1
echo "initcheck exit $? (decode and shard tests)" | tee -a $S; tail -2 gpurun_out/initcheck.log | tee -a $S
echo "initcheck error sites:" >> $S
grep -A1 "Uninitialized\|Host API memory access error" gpurun_out/initcheck.log | grep " at \|access by\|Host API" | sed 's/^=========  *//' | cut -c1-160 | sort | uniq -c | sort -rn | head -12 | tee -a $S
