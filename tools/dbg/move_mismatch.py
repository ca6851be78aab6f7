import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _bmc_loader import load_pkg, load_synth
import util, oracle
oracle.build()
bmc, synth = load_pkg(), load_synth()
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
if mode == "exact": os.environ["BMC_CAPACITY_MODE"] = "exact"
n = 200_000
case = util.make_case(synth, "fixed_length", n, 16, dt=100.0, p_move=0.3, outlet=False)
case["props"][0][:] = np.float32(1.9995e-6)
kw = dict(allocation_factor=1.5, buffer_ratio=0.6)
if mode == "big": kw = dict(allocation_factor=4.0, buffer_ratio=0.6)
g = bmc.ParticleLoop("fixed_length", 1, 16, seed=case["seed"], **kw)
o = oracle.OracleLoop("fixed_length", 1, 16, seed=case["seed"], n_threads=4, **kw)
util.load_case(g, case); util.load_case(o, case)
for s in range(12):
    g.set_concentrations(util.conc_at(case, s)); g.cycle(case["dt"])
    o.set_concentrations(util.conc_at(case, s)); o.cycle(case["dt"])
    cg, co = g.counters(), o.counters()
    nn = co["n_used"]
    pg, po = g.get_particles(nn), o.get_particles(nn)
    dpos = int(np.sum(pg["position"][:nn] != po["position"][:nn]))
    print(s, "n", cg["n_used"], co["n_used"], "move", cg["events"]["Move"], co["events"]["Move"], "new", cg["events"]["NewParticle"], co["events"]["NewParticle"],
          "ovf", cg["events"]["Overflow"], co["events"]["Overflow"], "cap", cg["capacity"], co["capacity"], "phys", cg["physical_capacity"], "dpos", dpos,
          "first", (np.nonzero(pg["position"][:nn] != po["position"][:nn])[0][:8]).tolist())
