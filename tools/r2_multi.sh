#!/bin/bash
# N GPUs of one box: the multi-process tests (N >= 2), then bench.py at N (fused peer exchange) for the headline and the c2 shard
T=${TAG:-r2multi}
N=${NG:-2}
( timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_p2p_allreduce_gpu.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${T}_tests.log 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps $2 --warmup 5 ${@:3}; }
run 29531 50 > gpurun_out/${T}_ns.json 2> gpurun_out/${T}.err
run 29533 500 --workload c2 > gpurun_out/${T}_c2.json 2>> gpurun_out/${T}.err
[ -n "$WITH_NCCL" ] && BMC_P2P=0 run 29532 50 > gpurun_out/${T}_ns_nccl.json 2>> gpurun_out/${T}.err
cat gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}.err
for f in gpurun_out/${T}_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d.get("collective_check"), d.get("parallelism","")[-60:])
except Exception as e: print("ERR", e, open("$f").read()[-300:])
PY
done
