"""per-step kernel time of the c3 workload (division + exits + compactions): which steps are slow"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from _bmc_loader import load_pkg, load_synth
pkg, synth = load_pkg(), load_synth()
wl = "c3"
model, n_comp, n, dt, near, p_exit = bench.WORKLOADS[wl]
n = int(os.environ.get("N", n))
p_exit = bench.P_EXIT.get(wl, p_exit)
fm, flows, conc = bench.build_case(synth, model, n_comp, dt, p_exit)
loop = pkg.ParticleLoop(model, 1, n_comp, **bench.LOOP_KW.get(wl, {}))
loop.init_particles(n, uniform_position=True)
props, pos = synth.make_population(model, n, n_comp, seed=11, near_division=near)
loop.set_particles(props, pos); del props, pos
loop.reserve(int(2.4 * n))
loop.set_weight(1e3)
bench.setup_loop(loop, fm, flows, conc, n_comp)
loop.profile_enable(1)
prev = loop.counters()
for s in range(24):
    loop.cycle(dt); loop.sync()
    ms, k = loop.profile_read()
    c = loop.counters()
    print(s, "kernel us %.1f" % (ms * 1e3), "n", c["n_used"], "new", c["total_new"] - prev["total_new"], "out", c["total_out"] - prev["total_out"],
          "inactive", c["n_inactive"], "compactions", c["n_compactions"] - prev["n_compactions"])
    prev = c
