#!/bin/bash
# A/B of two builds of the library (lib/ vs lib/alt/): kernel times under ncu, then the decode-only bench of each
mkdir -p gpurun_out
for v in main alt main alt; do
  if [ $v = alt ]; then export M3S_LIB_PATH=$PWD/mp3-steganography-lib_b200/lib/alt/libmp3stego_b200.so; else unset M3S_LIB_PATH; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_hybrid' -s 2 -c 3 --csv --log-file gpurun_out/ab_$v.csv \
      python bench.py --files 64 --steps 1 --warmup 1 --no-encode --no-extras > /dev/null 2>&1
  echo "== $v"; grep -v "^==" gpurun_out/ab_$v.csv | tail -6 | cut -d, -f5,12- | cut -c1-200
done
for v in main alt; do
  if [ $v = alt ]; then export M3S_LIB_PATH=$PWD/mp3-steganography-lib_b200/lib/alt/libmp3stego_b200.so; else unset M3S_LIB_PATH; fi
  timeout 600 python bench.py --steps 3 --warmup 2 --no-encode --no-extras > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$v.json')); print('$v', 'value %.4g ms %.1f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done
