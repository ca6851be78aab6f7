#!/bin/bash
# round 2 profile capture: launch list of the default bench command + ncu --set full of the step kernel per workload
T=${TAG:-r2}
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
for w in ns c2 c3 c4 fl_ns sa_ns; do
  timeout 1500 $NCU --set full --import-source on -k regex:cycle_kernel -s 6 -c 1 -f -o gpurun_out/${T}_cycle_$w python bench.py --workload $w --steps 8 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_ncu_$w.log 2>&1
done
timeout 1500 $NCU --set full --import-source on -k regex:cycle_kernel -s 6 -c 1 -f -o gpurun_out/${T}_cycle_c2_eager python bench.py --workload c2 --eager --steps 8 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_ncu_c2_eager.log 2>&1
timeout 1500 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum -k regex:cycle_kernel -s 6 -c 1 --csv --log-file gpurun_out/${T}_ncu_c5.csv python bench.py --workload c5 --steps 8 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_ncu_c5.log 2>&1
ls -la gpurun_out/${T}_*
