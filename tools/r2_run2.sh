#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2b}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/${T}_tests.log 2>&1
for v in v4b4 v4b3 v4b2; do
  BMC_VARIANT=$v timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_c2_$v.log 2>&1
  BMC_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --particles 125000000 > gpurun_out/${T}_bench_ns_$v.log 2>&1
done
for v in v4b4 v4b3; do
  BMC_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:cycle_kernel -s 6 -c 2 --csv --log-file gpurun_out/${T}_ncu_c2_$v.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_c2_$v.log 2>&1
done
tail -3 gpurun_out/${T}_tests.log; for f in gpurun_out/${T}_bench_*.log; do echo $f; python - <<PY
import json,sys
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("ERR", e, open("$f").read()[-500:])
PY
done
