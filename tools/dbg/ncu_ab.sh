#!/bin/bash
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for L in lib_old.so libbmc_b200.so; do
  BMC_LIB=$PWD/biocma-mcst_b200/$L ncu --metrics $M --clock-control none -k regex:cycle_kernel -s 6 -c 1 --csv --log-file gpurun_out/ab_$L.csv python bench.py --workload c5 --particles 40000000 --steps 8 --warmup 3 --no-configs --no-cpu-baseline > /dev/null 2>&1
  echo "== $L"; python - <<PY
import csv
for r in csv.reader(open("gpurun_out/ab_$L.csv")):
    if len(r) > 6 and r[0].isdigit(): print(r[-3], r[-1])
PY
done
