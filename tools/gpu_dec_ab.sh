#!/bin/bash
# decode parity tests with the current build, then the decode-only bench of the current build and of lib/alt/ (A/B on one box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py tests/test_batch_gpu.py tests/test_facade_gpu.py -m gpu -q -x > gpurun_out/tests_dec.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_dec.log
tail -3 gpurun_out/tests_dec.log
for v in main alt main alt; do
  if [ $v = alt ]; then export M3S_LIB_PATH=$PWD/mp3-steganography-lib_b200/lib/alt/libmp3stego_b200.so; else unset M3S_LIB_PATH; fi
  timeout 600 python bench.py --steps 3 --warmup 2 --no-encode --no-extras > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$v.json')); print('$v', 'value %.4g ms %.1f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['check'].get('parity_sampled'))"
done
