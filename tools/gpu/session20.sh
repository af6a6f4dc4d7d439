#!/bin/bash
O=gpurun_out/s20
mkdir -p $O
build/sort_plain_old | grep -E "cells|check" > $O/plain.txt
build/sort_plain | grep -E "cells|check" >> $O/plain.txt
cat $O/plain.txt
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for v in old new; do
  b=build/sort_plain; [ $v = old ] && b=build/sort_plain_old
  timeout 200 ncu --metrics $M --clock-control none -k regex:rs_sort_test -s 24 -c 1 --csv --log-file $O/ncu_$v.csv $b > /dev/null 2>&1
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:rs_sort_test -s 24 -c 1 -o $O/sort_$v $b > /dev/null 2>&1
done
grep -h "rs_sort" $O/ncu_old.csv | awk -F'","' '{print $(NF-2), $NF}' | head -20
echo ==; grep -h "rs_sort" $O/ncu_new.csv | awk -F'","' '{print $(NF-2), $NF}' | head -20
