#!/usr/bin/env python
"""Per-step device times of the step kernel (CUDA events through bmc_profile_*): python tools/steptimes.py [particles] [steps] [workload]"""
import os, sys, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _bmc_loader import load_pkg, load_synth
import util
pkg, synth = load_pkg(), load_synth()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ncomp = int(sys.argv[3]) if len(sys.argv) > 3 else 500
fm = synth.make_flowmap(ncomp, 0.1, p_move=0.01, seed=2024)
g = pkg.ParticleLoop("monod", 1, ncomp)
g.init_particles(n, True)
g.set_weight(1e3)
g.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
o = ncomp - 1
g.set_leaving_flows([(o, 1e-3 * fm["volumes"][o] / 0.1, fm["volumes"][o])])
g.set_concentrations(np.full(ncomp, 5.0))
g.profile_enable(True)
times = []
sync_each = os.environ.get("SYNC_EACH", "0") == "1"
for s in range(steps):
    g.cycle(0.1)
    if sync_each:
        g.sync()
        if os.environ.get('SLEEP_MS'):
            import time; time.sleep(float(os.environ['SLEEP_MS']) * 1e-3)
    if (s + 1) % 10 == 0 or sync_each:
        ms, k = g.profile_read()
        times.append((s + 1, ms / max(k, 1) * 1e3))
print("variant", os.environ.get("BMC_VARIANT"), "n", n, "sync_each", sync_each, "avg kernel us per block of steps:", [(s, round(t, 1)) for s, t in times[:12]])
