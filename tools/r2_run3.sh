#!/bin/bash
T=r2c
python tools/timeline.py > gpurun_out/${T}_timeline.log 2>&1
BMC_VARIANT=v4b4 python tools/steptimes.py 125000000 80 > gpurun_out/${T}_steptimes_b4.log 2>&1
BMC_VARIANT=v4b4 python tools/steptimes.py 100000000 80 > gpurun_out/${T}_steptimes_b4_1e8.log 2>&1
for w in c3 c4 c5 sa fl; do
  BMC_VARIANT=v4b4 timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${w}_b4.log 2>&1
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${w}_def.log 2>&1
done
cat gpurun_out/${T}_timeline.log gpurun_out/${T}_steptimes*.log
for f in gpurun_out/${T}_bench_*.log; do echo $f; python - <<PY
import json,sys
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("ERR", e, open("$f").read()[-500:])
PY
done
