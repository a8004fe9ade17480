#!/bin/bash
# ncu --set full of the analysis kernel under each variant (1,000 clips x 100 frames per launch)
mkdir -p gpurun_out
for v in "${@:-1}"; do
  case $v in direct) export M3S_ENC_ANALYSIS_DIRECT=1; unset M3S_ENC_FOLD_CFG;; *) unset M3S_ENC_ANALYSIS_DIRECT; export M3S_ENC_FOLD_CFG=$v;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_enc_analysis' -s 2 -c 1 -f -o gpurun_out/prof_ana_$v \
      python bench.py --files 1000 --frames ${FRAMES:-100} --steps 1 --warmup 1 --no-extras > gpurun_out/ncu_ana_$v.log 2>&1
  python tools/ncu_pick.py gpurun_out/prof_ana_$v.ncu-rep gpurun_out/ana_${v}_summary.csv
done
