#!/bin/bash
# ncu counters of the likelihood tail's kernels (one capture per launch, first 6 launches): duration, instructions, issue
# utilisation, stall reasons.  Usage (GPU box): bash tools/ncu_plik.sh <out.csv>
ncu --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:plik -c 6 --csv --log-file "$1" python tools/time_likelihood.py
