"""Liquid phase on the device ("next" row 1): bmc_liquid_step = update_feed (scalar part) +
ScalarSimulation::performStep + clearContribution (implScalar.cpp:251-266,
simulation.model.cpp:55-154), and the fully device-resident time loop built from it."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _feed_terms(feeds, ns, nc):
    src = np.zeros(ns * nc); sink = np.zeros(nc)
    for f in feeds:
        src[f["species"] + ns * f["input_position"]] += f["flow"] * f["concentration"]
        if f.get("output_position") is not None and f.get("first_of_feed", 1):
            sink[f["output_position"]] += f["flow"]
    return src, sink


@pytest.mark.parametrize("n_comp,ns,model", [(1, 1, "monod"), (64, 1, "monod"), (500, 2, "simple_acetate")])
def test_liquid_step_bit_exact(bmc, orc, synth, n_comp, ns, model):
    fm = synth.make_flowmap(n_comp, 0.1, p_move=0.05)
    rng = np.random.default_rng(3)
    C0 = rng.random(ns * n_comp) + 0.5
    g = bmc.ParticleLoop(model, ns, n_comp)
    g.domain_update(fm["volumes"], fm["neighbors"] if n_comp > 1 else None, fm["out_flows"], fm["cdf"] if n_comp > 1 else None)
    g.liquid_set_transition(fm["coo"])
    g.set_concentrations(C0)
    feeds = [dict(species=0, input_position=0, flow=3e-5, concentration=5.0, output_position=n_comp - 1)]
    if ns > 1:
        feeds.append(dict(species=1, input_position=0, flow=3e-5, concentration=0.7, output_position=n_comp - 1, first_of_feed=0))
    g.liquid_set_feeds(feeds)
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    volx = np.repeat(vol, ns)
    C = C0.copy(); mass = C * volx
    for step in range(40):
        g.liquid_step(0.1)
        src, sink = _feed_terms(feeds, ns, n_comp)
        orc.ode_step(C, mass, vol, sink, src, fm["coo"], 0.1)
    got = g.get_concentrations()
    assert np.array_equal(got.view(np.uint64), C.view(np.uint64)), np.max(np.abs(got - C))
    # closed tank: mass conserved up to the feed/outlet balance
    assert np.all(got > 0)


def test_device_resident_time_loop_matches_oracle_loop(bmc, orc, synth):
    # every step: feed -> ODE -> clear -> cycleProcess, nothing crosses PCIe except the final read-back
    n_comp, n, dt = 100, 150_000, 5.0
    case = util.make_case(synth, "monod", n, n_comp, dt=dt, near_division=0.6, p_move=0.2, p_exit=0.0, outlet=False)
    fm = case["fm"]
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    q = 0.02 * vol[n_comp - 1] / dt
    feeds = [dict(species=0, input_position=0, flow=q, concentration=8.0, output_position=n_comp - 1)]
    C0 = np.full(n_comp, 2.0)
    g = bmc.ParticleLoop("monod", 1, n_comp, seed=11)
    o = orc.OracleLoop("monod", 1, n_comp, seed=11, n_threads=4)
    for L in (g, o):
        L.set_particles(case["props"], case["pos"]); L.set_weight(case["weight"] * 2e3)
        L.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    g.liquid_set_transition(fm["coo"]); g.set_concentrations(C0); g.liquid_set_feeds(feeds)
    o.set_leaving_flows([(n_comp - 1, q, vol[n_comp - 1])])
    C = C0.copy(); mass = C * vol; sources = np.zeros(n_comp)
    traj_g, traj_o = [], []
    for step in range(30):
        g.liquid_step(dt); g.cycle(dt)
        src, sink = _feed_terms(feeds, 1, n_comp)
        sources += src
        orc.ode_step(C, mass, vol, sink, sources, fm["coo"], dt)
        o.set_concentrations(C); o.cycle(dt)
        sources = o.get_sources().copy()
        if step % 5 == 4:
            traj_g.append(g.get_concentrations()); traj_o.append(C.copy())
    np.testing.assert_allclose(np.array(traj_g), np.array(traj_o), rtol=1e-9, atol=0)  # concentration trajectory
    assert np.max(np.abs(np.array(traj_o)[-1] - 2.0)) > 1e-3                            # uptake + feed did change it
    util.assert_counters_equal(g.counters(), o.counters())
    assert np.array_equal(g.repartition(), o.repartition())                             # occupancy
    assert g.counters()["total_out"] > 0 and g.counters()["total_new"] > 0
