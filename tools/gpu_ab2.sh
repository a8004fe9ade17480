#!/bin/bash
# A/B of three builds (lib/ = channel count dispatched inside one kernel, lib/alt = run-time channel count, lib/alt2 = one kernel
# per channel count + light probes in the encoder): hybrid kernel times under ncu; alt2: encode parity tests + bench with the
# certified-bounds shortcut on / off
mkdir -p gpurun_out
L=$PWD/mp3-steganography-lib_b200/lib
for v in main alt alt2 main alt alt2; do
  if [ $v = main ]; then unset M3S_LIB_PATH; else export M3S_LIB_PATH=$L/$v/libmp3stego_b200.so; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_hybrid' -s 2 -c 1 --csv --log-file gpurun_out/ab_$v.csv \
      python bench.py --files 64 --steps 1 --warmup 1 --no-encode --no-extras > /dev/null 2>&1
  echo "== $v $(grep -v '^==' gpurun_out/ab_$v.csv | tail -2 | awk -F, '{print $NF}' | tr '\n' ' ')"
done
export M3S_LIB_PATH=$L/alt2/libmp3stego_b200.so
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_alt2.log 2>&1; echo "pytest alt2 exit $?"; tail -4 gpurun_out/tests_alt2.log
for light in 1 0; do
  export M3S_PROBE_LIGHT=$light
  timeout 900 python bench.py --steps 3 --warmup 2 --no-extras > gpurun_out/bench_light$light.json 2> gpurun_out/bench_light$light.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_light$light.json')); e=d['encode_hide']; print('light=$light decode %.4g (%.1f ms) encode %.4g (%.1f ms) e2e %.4g' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['e2e']['value']), {k: round(x,1) for k,x in e['roofline']['kernel_ms_per_step'].items()}, d['check'])"
done
