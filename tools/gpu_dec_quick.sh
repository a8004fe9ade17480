#!/bin/bash
# decode parity tests + the decode-only bench line (kernel ms per step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_decode.py tests/test_shard_gpu.py -m gpu -q -x > gpurun_out/tests_dec.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_dec.log
tail -3 gpurun_out/tests_dec.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-encode --no-extras > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_dec.json')); print('value %.4g ms %.1f e2e %.4g (%.1f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['check'].get('parity_sampled'))"
