#!/bin/bash
# Diagnostic: what host does this GPU box have (cores, NUMA, memory bandwidth with streaming stores, PCIe rates)?
echo "== lscpu"; lscpu | egrep -i "model name|^CPU\(s\)|thread|core|socket|numa|L3|L2 cache|MHz" 
echo "== affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0)))"
echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; numactl -H 2>/dev/null | head -20
echo "== hugepages"; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>/dev/null; grep -i huge /proc/meminfo
echo "== mem"; free -g | head -2
echo "== gpu topo"; nvidia-smi topo -m 2>/dev/null | head -20; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
cd "$(dirname "$0")/microbench" && gcc -O2 -mavx2 -fopenmp -o host_bw host_bw.c && for t in 4 6 8 10 12 14 16 20 24 32; do [ $t -le $(nproc) ] && ./host_bw $t | head -2; done
