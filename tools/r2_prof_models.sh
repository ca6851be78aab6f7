#!/bin/bash
# ncu --set full of the step kernel for the other built-in models at the north-star population
T=${TAG:-r2m}
NCU="ncu --clock-control none"
for w in ${WL:-fl sa}; do
  timeout 900 $NCU --set full --import-source on -k regex:cycle_kernel -s 6 -c 1 -f -o gpurun_out/${T}_cycle_${w} python bench.py --workload $w --particles 125000000 --steps 8 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_ncu_$w.log 2>&1
done
ls -la gpurun_out/${T}_*
