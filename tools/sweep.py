#!/usr/bin/env python
"""Tuning sweep over builds (BMC_LIB), kernel variants (BMC_VARIANT) and the L2 prefetch
distance (BMC_PREFETCH): runs bench.py for each combination and prints one line each."""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = os.environ.get("SWEEP_LIBS", "libbmc_b200.so").split(",")
variants = os.environ.get("SWEEP_VARIANTS", "v4b1").split(",")
prefetch = os.environ.get("SWEEP_PREFETCH", "0").split(",")
extra = sys.argv[1:]
for lib, var, pf in itertools.product(libs, variants, prefetch):
    env = dict(os.environ, BMC_LIB=os.path.join(ROOT, "biocma-mcst_b200", lib), BMC_VARIANT=var, BMC_PREFETCH=pf)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "300", "--warmup", "10", "--no-cpu-baseline"] + extra,
                       capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{lib:28s} {var:5s} pf={pf}  step={d['ms_per_step']*1e3:7.1f}us kernel={d['roofline']['kernel_ms']*1e3:7.1f}us "
              f"frac={d['roofline']['frac']:.3f} value={d['value']:.3e} e2e={d['e2e']['value']:.3e}", flush=True)
    except Exception as e:
        print(lib, var, pf, "FAILED", r.stderr[-400:], flush=True)
