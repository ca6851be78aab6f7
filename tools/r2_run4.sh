#!/bin/bash
T=${TAG:-r2d}
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${T}_tests.log 2>&1
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_c2.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --particles 125000000 > gpurun_out/${T}_bench_ns.log 2>&1
for w in sa; do for v in v4b4 v4b3; do
  BMC_VARIANT=$v timeout 600 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_${w}_$v.log 2>&1
done; done
cat gpurun_out/${T}_tests.log
for f in gpurun_out/${T}_bench_*.log; do echo $f; python - <<PY
import json,sys
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("ERR", e, open("$f").read()[-500:])
PY
done
