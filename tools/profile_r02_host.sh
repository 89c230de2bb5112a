#!/bin/bash
# Round-2 profiling pass, second half (host-buffer path + single-step kernel); see tools/profile_r02.sh.  Bounded: every
# profiler run under its own timeout.
set -x
O=gpurun_out
H="python tools/host_loop.py 65536 30"
CUDA_LAUNCH_BLOCKING=1 timeout 120 $H > $O/r02_host_blocking.log 2>&1 || { echo "host loop under CUDA_LAUNCH_BLOCKING failed"; exit 1; }
timeout 120 $H > $O/r02_host_plain.log 2>&1 || exit 1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_65k_host.csv $H > $O/r02_ncu_hl.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"spl_push_kernel|spl_step_kernel" -s 20 -c 2 -f -o $O/r02_prof_host_65k $H > $O/r02_ncu_hf.log 2>&1
S="python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-lockstep --skip-configs --mode lockstep --no-graph --rollout 8"
timeout 120 $S > $O/r02_step_plain.log 2>&1 || exit 1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_65k_lockstep.csv $S > $O/r02_ncu_sl.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:spl_step_kernel -s 20 -c 1 -f -o $O/r02_prof_step_65k $S > $O/r02_ncu_sf.log 2>&1
tail -2 $O/r02_ncu_hl.log $O/r02_ncu_hf.log $O/r02_ncu_sl.log $O/r02_ncu_sf.log
