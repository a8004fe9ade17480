#!/bin/bash
# ncu --set full of the parallel rate loop kernels (1,000 clips x 200 frames per launch)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_enc_probe|k_enc_emit|k_enc_resolve' -s 3 -c 3 -f -o gpurun_out/prof_probe \
    python tools/enc_ab.py 1000 200 default > gpurun_out/ncu_probe.log 2>&1; echo "ncu exit $?"
tail -5 gpurun_out/ncu_probe.log
ls -la gpurun_out
