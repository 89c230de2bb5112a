#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun).  Every ncu command runs only after the same command has
# exited 0 without ncu; numbers printed under ncu are never bench values.  Every profiler run is under its own `timeout`: a
# profiler serialises launches, and a kernel that waits for host threads (the push kernel's flow control) must never be able
# to hold the box until gpurun's limit.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-lockstep --skip-configs"
$B > $O/r02_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_65k_rollout.csv $B > $O/r02_ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spl_rollout_kernel -s 3 -c 1 -f -o $O/r02_prof_rollout_65k $B > $O/r02_ncu_f.log 2>&1
H="python tools/host_loop.py 65536 30"
$H > $O/r02_host_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_65k_host.csv $H > $O/r02_ncu_hl.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"spl_push_kernel|spl_step_kernel" -s 20 -c 2 -f -o $O/r02_prof_host_65k $H > $O/r02_ncu_hf.log 2>&1
S="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-lockstep --skip-configs --mode lockstep --no-graph --rollout 8"
$S > $O/r02_step_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_65k_lockstep.csv $S > $O/r02_ncu_sl.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spl_step_kernel -s 20 -c 1 -f -o $O/r02_prof_step_65k $S > $O/r02_ncu_sf.log 2>&1
ls -la $O | grep r02
