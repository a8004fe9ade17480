#!/bin/bash
# round 2, pass e (N GPUs of one box): the bench under torchrun, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
{ echo "== numa"; ls /sys/devices/system/node/ | grep node; for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done; nproc; head -3 /proc/meminfo; } >> gpurun_out/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
grep "rank 0\]" gpurun_out/bench_n$N.err | tail -8
python - $N <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_n%s.json" % sys.argv[1]))
    print("decode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["roof_frac"]))
    print(" host roof", {k: v for k, v in d["e2e"]["host_roof_gbs"].items() if k != "how"})
    e = d.get("encode_hide")
    if e: print("encode value %.4g e2e %.4g ms %.1f e2e_ms %.1f roof_frac %.3f" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["e2e"]["ms_per_step"], e["e2e"]["roof_frac"]))
    print(" check", d["check"]["parity_sampled"]); c = d.get("cfg5")
    if c: print(" cfg5 value %.4g e2e %.4g eff %.3f e2e_eff %.3f parity %s long %s" % (c["value"], c["e2e"]["value"], c["efficiency"], c["e2e_efficiency"], c["parity_ok"], c["long_file"]))
    print(" composite", d.get("composite"))
except Exception as ex:
    print("bench parse failed:", ex)
PY
tail -5 gpurun_out/bench_n$N.err | cut -c1-300
