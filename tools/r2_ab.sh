#!/bin/bash
# quick A/B of the step kernel: WL = workloads, ENVS = ';'-separated environment settings (each "NAME=V NAME2=V2" or "-")
T=${TAG:-r2ab}
IFS=';' read -ra ES <<< "${ENVS:--}"
i=0
for e in "${ES[@]}"; do
  for w in ${WL:-ns c2 fl_ns sa_ns}; do
    steps=30; [ $w = c2 ] && steps=500; [ $w = c5 ] && steps=8
    ( [ "$e" != "-" ] && export $e; timeout 300 python bench.py --workload $w --steps $steps --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/${T}_${i}_${w}.json 2> gpurun_out/${T}_${i}_${w}.err )
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_${i}_${w}.json").read().strip().splitlines()[-1])
    print("[$e] $w", round(d["ms_per_step"],5), "ms  frac", round(d["roofline"]["frac"],4), " e2e", round(d["e2e"]["ms_per_step"],5), d["roofline"]["kernel"][13:50])
except Exception as ex: print("[$e] $w ERR", ex, open("gpurun_out/${T}_${i}_${w}.err").read()[-500:])
PY
  done
  i=$((i+1))
done
