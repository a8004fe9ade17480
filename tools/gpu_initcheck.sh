#!/bin/bash
# compute-sanitizer initcheck over the decode parity tests, summarised by error site
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool initcheck --print-limit 200000 --error-exitcode 9 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py -m gpu -x -q > gpurun_out/initcheck_full.log 2>&1
echo "initcheck exit $?"; tail -3 gpurun_out/initcheck_full.log | cut -c1-200
grep " at " gpurun_out/initcheck_full.log | sed 's/^=========     at //' | sort | uniq -c | sort -rn | head -20 > gpurun_out/initcheck_sites.txt; cat gpurun_out/initcheck_sites.txt
for k in k_huff k_hybrid k_strip k_sideinfo k_fscan k_walk; do echo "$k: $(grep -c " at .*$k" gpurun_out/initcheck_full.log)"; done
grep -A14 "at .*k_huff\|at .*k_hybrid" gpurun_out/initcheck_full.log | head -40 > gpurun_out/initcheck_samples.txt
rm -f gpurun_out/initcheck_full.log
