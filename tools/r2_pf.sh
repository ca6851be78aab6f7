#!/bin/bash
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 
for pfv in 1 0; do for w in ns c2 c3 fl sa; do
  st=40; [ $w = c2 ] && st=400; [ $w = fl ] && st=400; [ $w = sa ] && st=200
  BMC_PREFETCH=$pfv python bench.py --workload $w --steps $st --warmup 5 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('prefetch $pfv $w', round(d['ms_per_step']*1e3,1), 'us', '%.3e' % d['value'], round(d['roofline']['frac'],3), 'e2e %.3e' % d['e2e']['value'])"
done; done
