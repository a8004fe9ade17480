#!/bin/bash
# round 2, pass g: GPU tests, configs[4] alone (long-file walk), decode-only bench, compute-sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 3 --warmup 2 --only-cfg5 > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err; echo "cfg5 exit $?"; cut -c1-1800 gpurun_out/cfg5_n1.json
timeout 600 python bench.py --steps 3 --warmup 2 --no-encode --no-extras > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err
python -c "
import json; d=json.load(open('gpurun_out/bench_dec.json')); print('decode %.4g (%.1f ms) e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items()}, d['check']['parity_sampled'])"
bash tools/gpu_sanitize.sh
