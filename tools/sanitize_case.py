"""Small scenarios of every kernel family for compute-sanitizer (tools/sanitize.sh): each is checked against the oracle
as the parity tests do, so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from _bmc_loader import load_pkg, load_synth  # noqa: E402
import oracle  # noqa: E402
import util  # noqa: E402

pkg, synth = load_pkg(), load_synth()


def run(model, n, n_comp, steps, eager=False, **kw):
    ns = 2 if model == "simple_acetate" else 1
    case = util.make_case(synth, model, n, n_comp, **kw)
    if model == "simple_acetate":
        case["props"][1, :] = 1.0  # no division: its division draws through libdevice (last-bit differences)
    g = pkg.ParticleLoop(model, ns, n_comp)
    o = oracle.OracleLoop(model, ns, n_comp, n_threads=2)
    util.load_case(g, case); util.load_case(o, case)
    if eager:  # caller-supplied ages: the eager-age kernel
        ages = np.linspace(0.0, 5.0, n).astype(np.float32)
        for loop in (g, o):
            loop.set_particles(case["props"], case["pos"], None, ages, ages[::-1].copy())
        assert not g.kernel_config()["stamped_ages"]
    sg = util.run_steps(g, case, steps, collect=True); so = util.run_steps(o, case, steps, collect=True)
    cg, co = g.counters(), o.counters()
    util.assert_counters_equal(cg, co)
    n_used = co["n_used"]
    util.assert_state_equal(g.get_particles(n_used), o.get_particles(n_used), n_used)
    for a, b in zip(sg, so):
        assert np.max(np.abs(a - b)) <= 1e-9 * (np.max(np.abs(b)) + 1e-300)
    g.compact()
    blob = g.checkpoint()
    g.restore(blob)
    g.get_properties()
    print(f"case ok: {model} n={n} comp={n_comp} eager={eager} n_used={n_used} new={cg['total_new']} out={cg['total_out']} "
          f"compactions={cg['n_compactions']} ckpt={len(blob)}")


run("monod", 20_000, 500, 4, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
run("monod", 5_001, 1, 3, dt=20.0, near_division=0.8)                       # 0D: register accumulation, ragged tail
run("fixed_length", 9_000, 37, 4, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
run("simple_acetate", 9_000, 37, 4, dt=20.0, p_move=0.3, p_exit=0.3)
run("monod", 9_000, 37, 4, eager=True, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
run("monod", 6_000, 9_000, 2, dt=20.0, p_move=0.3, p_exit=0.3)             # table and bins too large for shared memory
print("all cases ok")
