#!/bin/bash
# 2 GPUs: p2p test on both, then bench at N=2 (fused peer exchange) vs NCCL, ns and c2 sizes
T=${TAG:-r2g}
( timeout 900 python -m pytest tests/test_p2p_allreduce_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${T}_tests.log 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps $2 --warmup 5 ${@:3}; }
run 29521 50 > gpurun_out/${T}_n2_ns.json 2> gpurun_out/${T}_n2.err
BMC_P2P=0 run 29522 50 > gpurun_out/${T}_n2_ns_nccl.json 2>> gpurun_out/${T}_n2.err
run 29523 500 --workload c2 > gpurun_out/${T}_n2_c2.json 2>> gpurun_out/${T}_n2.err
BMC_P2P=0 run 29524 500 --workload c2 > gpurun_out/${T}_n2_c2_nccl.json 2>> gpurun_out/${T}_n2.err
python bench.py --steps 500 --warmup 5 --workload c2 --no-cpu-baseline > gpurun_out/${T}_n1_c2.json 2>> gpurun_out/${T}_n2.err
python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/${T}_n1_ns.json 2>> gpurun_out/${T}_n2.err
cat gpurun_out/${T}_tests.log; tail -5 gpurun_out/${T}_n2.err
for f in gpurun_out/${T}_n*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d.get("collective_check"), d.get("parallelism","")[-60:])
except Exception as e: print("ERR", e, open("$f").read()[-300:])
PY
done
