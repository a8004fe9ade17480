#!/bin/bash
# ncu --set full capture of selected kernels.  usage: KREGEX='k_hybrid' SKIP=2 COUNT=1 tools/gpu_prof.sh [bench args]
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_hybrid}" -s ${SKIP:-2} -c ${COUNT:-1} -o gpurun_out/prof_${TAG:-k} \
    python bench.py --files 16 --wave 16 --steps 1 --warmup 1 "$@" > gpurun_out/ncu_prof.log 2>&1
tail -3 gpurun_out/ncu_prof.log; ls -la gpurun_out/*.ncu-rep
