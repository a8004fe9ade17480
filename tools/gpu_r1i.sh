#!/bin/bash
# GPU pass r1i: parity of the parallel rate loop (probe / resolve / emit), then A/B timing of the encoder pipeline variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_encode.py -m gpu -x -q > gpurun_out/tests_enc.log 2>&1; echo "enc pytest exit $?" | tee -a gpurun_out/tests_enc.log
tail -30 gpurun_out/tests_enc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "all pytest exit $?" | tee -a gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 900 python tools/enc_ab.py 1000 1378 > gpurun_out/enc_ab.log 2>&1; echo "enc_ab exit $?"
cat gpurun_out/enc_ab.log | tail -12
