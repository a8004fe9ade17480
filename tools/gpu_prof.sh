#!/bin/bash
# ncu --set full capture of selected kernels.
# usage: KREGEX='k_hybrid' SKIP=2 COUNT=1 TAG=hyb tools/gpu_prof.sh [bench args]   (default bench args: 16 files, decode only)
mkdir -p gpurun_out
ARGS="$@"; [ -z "$ARGS" ] && ARGS="--files 16 --wave 16 --steps 1 --warmup 1 --no-encode"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_hybrid}" -s ${SKIP:-2} -c ${COUNT:-1} -o gpurun_out/prof_${TAG:-k} -f \
    python bench.py $ARGS > gpurun_out/ncu_prof_${TAG:-k}.log 2>&1
tail -3 gpurun_out/ncu_prof_${TAG:-k}.log; ls -la gpurun_out/*.ncu-rep
