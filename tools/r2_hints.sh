#!/bin/bash
for h in 0 1 2 3; do
  L=biocma-mcst_b200/lib_h$h.so; [ $h = 0 ] && L=biocma-mcst_b200/libbmc_b200.so
  for w in ns c2; do
    st=40; [ $w = c2 ] && st=400
    BMC_LIB=$PWD/$L python bench.py --workload $w --steps $st --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2i_${w}_h$h.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_${w}_h$h.json").read().strip().splitlines()[-1])
print("hints $h $w", round(d["ms_per_step"]*1e3,1), "us", "%.3e" % d["value"], round(d["roofline"]["frac"],3))
PY
  done
done
