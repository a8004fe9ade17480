#!/bin/bash
# host topology probe of the GPU box: NUMA nodes, GPU <-> node affinity, cores, memory (profiles/r2a_topology.txt)
mkdir -p gpurun_out
{
echo "== nvidia-smi topo -m"; nvidia-smi topo -m
echo "== lscpu"; lscpu | head -40
echo "== numa nodes"; ls /sys/devices/system/node/ 2>/dev/null
for n in /sys/devices/system/node/node*; do echo "$n cpulist $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
echo "== nvidia-smi -q pci"; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv
echo "== affinity of this shell"; taskset -p $$; nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
echo "== meminfo"; head -5 /proc/meminfo
which numactl; python -c "import os; print('sched_getaffinity', len(os.sched_getaffinity(0)))"
} > gpurun_out/topology.txt 2>&1
tail -60 gpurun_out/topology.txt
