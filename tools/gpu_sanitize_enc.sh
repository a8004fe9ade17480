#!/bin/bash
# compute-sanitizer over the encode parity tests (folded analysis kernel, device-written frame -> clip map): memcheck, racecheck, initcheck
mkdir -p gpurun_out
S=gpurun_out/sanitizer_enc_summary.txt
echo "compute-sanitizer (CUDA 12.9) over the encode GPU tests on a B200 (tools/gpu_sanitize_enc.sh)" > $S
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_encode.py tests/test_batch_gpu.py -m gpu -x -q > gpurun_out/memcheck_enc.log 2>&1
echo "memcheck exit $? (encode parity + batch composites)" | tee -a $S; tail -2 gpurun_out/memcheck_enc.log | tee -a $S
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_parity_encode.py::test_batch_vs_oracle" \
    "tests/test_parity_encode.py::test_quiet_silent_and_loud" "tests/test_parity_encode.py::test_chunked_equals_single" "tests/test_parity_encode.py::test_other_sample_rates" -m gpu -x -q > gpurun_out/racecheck_enc.log 2>&1
echo "racecheck exit $? (four encode tests)" | tee -a $S; tail -2 gpurun_out/racecheck_enc.log | tee -a $S
grep -c "Race reported\|hazard" gpurun_out/racecheck_enc.log | tee -a $S
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest "tests/test_parity_encode.py::test_batch_vs_oracle" "tests/test_parity_encode.py::test_chunked_equals_single" -m gpu -x -q > gpurun_out/initcheck_enc.log 2>&1
echo "initcheck exit $? (two encode tests)" | tee -a $S; tail -2 gpurun_out/initcheck_enc.log | tee -a $S
grep " at " gpurun_out/initcheck_enc.log | sed 's/^=========     at //' | cut -c1-120 | sort | uniq -c | sort -rn | head -8 | tee -a $S
