#!/bin/bash
# the other built-in models at the north-star population (1.25e8 particles per GPU): default and alternative block size
T=${TAG:-r2m}
for w in ${WL:-fl sa}; do
  for v in ${VARIANTS:-def v4b3}; do
    BMC_VARIANT=$v timeout 300 python bench.py --workload $w --particles 125000000 --steps 30 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/${T}_${w}_${v}.json 2> gpurun_out/${T}_${w}_${v}.err
  done
done
for f in gpurun_out/${T}_*_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["bytes_per_particle"], d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e, open("$f").read()[-300:], open("$f".replace(".json",".err")).read()[-600:])
PY
done
