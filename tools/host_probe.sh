#!/bin/bash
# host topology of a GPU box (diagnostic): cores, NUMA nodes, GPU <-> CPU affinity, PCIe link
echo "== lscpu"; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread|Core|L2|L3"
echo "== affinity"; python -c "import os; s=sorted(os.sched_getaffinity(0)); print(len(s), s)"
echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null; for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4, $5}')"; done
echo "== hugepages"; cat /sys/kernel/mm/transparent_hugepage/enabled; grep -i huge /proc/meminfo | head -4
echo "== gpu topo"; nvidia-smi topo -m 2>/dev/null | head -20
nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv
