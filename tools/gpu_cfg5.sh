#!/bin/bash
# configs[4] alone under torchrun on N GPUs (diagnostic)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --only-cfg5 > gpurun_out/cfg5_n$N.json 2> gpurun_out/cfg5_n$N.err; echo "cfg5 N=$N exit $?"
cat gpurun_out/cfg5_n$N.json | cut -c1-2500; grep -n "Error\|error" gpurun_out/cfg5_n$N.err | head -5
