#!/bin/bash
# round 2, pass k -- the full single-GPU pass behind profiles/r2k_* (folded analysis kernel, encode pipeline start-up, k_huff reader):
# same steps as tools/gpu_r2f.sh
exec bash tools/gpu_r2f.sh
