#!/bin/bash
# r2j: folded analysis kernel -- encode parity under several kernel shapes, then A/B of the shapes (kernel ms per pass)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_parity_encode.py tests/test_batch_gpu.py -m gpu -q -x > gpurun_out/tests_enc.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_enc.log
tail -4 gpurun_out/tests_enc.log
for g in ${PARITY_CFGS:-2 4 6}; do
  M3S_ENC_FOLD_CFG=$g timeout 600 python -m pytest tests/test_parity_encode.py -m gpu -q -x > gpurun_out/tests_enc_g$g.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_enc_g$g.log
  echo "cfg $g:"; tail -2 gpurun_out/tests_enc_g$g.log
done
args=""
for g in ${AB_CFGS:-0 1 2 3 4 5 6 7 8 9}; do args="$args env:M3S_ENC_FOLD_CFG=$g"; done
timeout 900 python tools/enc_ab.py 1000 1378 $args env:M3S_ENC_ANALYSIS_DIRECT=1 > gpurun_out/enc_ab.log 2>&1
cat gpurun_out/enc_ab.log | cut -c1-260
