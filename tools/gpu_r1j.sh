#!/bin/bash
# GPU pass: encode parity, A/B timing, ncu of the probe kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_parity_encode.py tests/test_batch_gpu.py tests/test_facade_gpu.py -m gpu -x -q > gpurun_out/tests_enc.log 2>&1; echo "enc pytest exit $?" | tee -a gpurun_out/tests_enc.log
tail -30 gpurun_out/tests_enc.log
timeout 900 python tools/enc_ab.py 1000 1378 ${AB_CONFIGS:-default serial} > gpurun_out/enc_ab.log 2>&1; echo "enc_ab exit $?"
tail -12 gpurun_out/enc_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_probe' -s 1 -c 1 -f -o gpurun_out/prof_probe \
    python tools/enc_ab.py 1000 200 default > gpurun_out/ncu_probe.log 2>&1; echo "ncu exit $?"
