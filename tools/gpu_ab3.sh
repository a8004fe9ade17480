#!/bin/bash
# GPU tests of the current build, then the bench (no extras) of the current build and of lib/alt (previous encoder / hybrid kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/tests_gpu.log
for v in main alt; do
  if [ $v = main ]; then unset M3S_LIB_PATH; else export M3S_LIB_PATH=$PWD/mp3-steganography-lib_b200/lib/$v/libmp3stego_b200.so; fi
  timeout 900 python bench.py --steps 3 --warmup 2 --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); e=d['encode_hide']; print('$v decode %.4g (%.1f ms) encode %.4g (%.1f ms) e2e %.4g' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['e2e']['value']), {k: round(x,1) for k,x in e['roofline']['kernel_ms_per_step'].items()}, {k: round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['check']['parity_sampled'])"
done
