#!/bin/bash
# round 2, final pass: tools/gpu_r2f.sh (tests, smoke, bench both arms, launch list, ncu --set full) + initcheck, on the shipped code
bash tools/gpu_r2f.sh
bash tools/gpu_initcheck.sh
