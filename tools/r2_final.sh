#!/bin/bash
# what the driver runs at round end: GPU suite, smoke, reference arm, default bench (timed)
T=${TAG:-r2fin}
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${T}_tests.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1
( time python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_ref.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${T}_bench.json") if l.startswith("{")][-1])
print("headline", d["ms_per_step"], d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
for k,v in d.get("configs",{}).items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("roofline",{}).get("frac"), v.get("roofline",{}).get("traffic"), v.get("error"))
r=json.loads([l for l in open("gpurun_out/${T}_ref.json") if l.startswith("{")][-1]); print("ref", r["value"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
PY
