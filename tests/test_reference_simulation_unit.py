"""The reference's OWN Simulation::SimulationUnit — simulation.cpp, simulation.model.cpp, simulation.getset.cpp,
scalar_init.cpp, feed_descriptor.cpp, implScalar.cpp, the kernels and the container — compiled where the sources lie
over the Kokkos / Eigen / rcmtool stand-ins (oracle/ref_sim.cpp, oracle/_ref/libbmc_ref_sim.so) and stepped by the body
of its main loop (apps/core/src/host_specific.cpp:281-291): update_feed, ode_step, advance, clearContribution,
cycleProcess.  Nothing of the time step is restated on the reference side.

Against it: the oracle's coupled loop (orc.ode_step + OracleLoop.cycle with the feed's outlet), i.e. the loop the CUDA
path is compared with bit for bit in tests/test_liquid_gpu.py::test_device_resident_time_loop_matches_oracle_loop.
Integer state (positions, statuses, container counters, event tallies) must be identical on every step; the
reference accumulates the source terms in float (ScatterView<float>), so sources agree to 5e-5 and the
concentration trajectory to 1e-4 relative (north_star: concentration trajectories within a stated tolerance)."""
import numpy as np
import pytest

import ref as refmod
import util

pytestmark = pytest.mark.skipif(not refmod.sim_available(), reason="oracle/_ref/libbmc_ref_sim.so not built (needs /root/reference)")


@pytest.mark.parametrize("model,steps", [("fixed_length", 36), ("monod", 24)])
def test_coupled_loop_of_the_reference_simulation_unit(orc, synth, model, steps):
    n_comp, n, dt = 40, 30_000, 5.0
    case = util.make_case(synth, model, n, n_comp, dt=dt, near_division=0.6, p_move=0.2, p_exit=0.0, outlet=False)
    fm = case["fm"]
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    q = 0.05 * vol[n_comp - 1] / dt                      # chemostat: feed into compartment 0, outlet in the last one
    feeds = [dict(species=0, input_position=0, flow=q, concentration=8.0, output_position=n_comp - 1)]
    C0 = np.full(n_comp, 2.0)
    w = case["weight"] * 2e3                             # enough biomass for the uptake to move the concentrations
    R = refmod.RefSim(model, 1, n_comp, vol, C0, feeds, seed=11)
    R.update_hydro(fm)
    R.set_particles(case["props"], case["pos"], w)
    o = orc.OracleLoop(model, 1, n_comp, seed=11, n_threads=2)
    o.set_particles(case["props"], case["pos"]); o.set_weight(w)
    o.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    o.set_leaving_flows([(n_comp - 1, q, vol[n_comp - 1])])   # update_feed: domain.set_leaving_flow(output, flow, volume)
    o.set_quirk_contrib_return(True)                     # the reference's contribution loop, bug for bug
    C = C0.copy(); mass = C * vol; mc = np.zeros(n_comp)
    worst_c = worst_s = 0.0
    for step in range(steps):
        R.step(dt)
        src = mc.copy(); src[0] += q * 8.0               # sources = Monte-Carlo contribution, then set_feed adds flow * concentration
        sink = np.zeros(n_comp); sink[n_comp - 1] += q
        orc.ode_step(C, mass, vol, sink, src, fm["coo"], dt)
        o.set_concentrations(C); o.cycle(dt)
        mc = o.get_sources().copy()
        cr, co = R.counters(), o.counters()
        util.assert_counters_equal(cr, co)
        worst_c = max(worst_c, float(np.max(np.abs(R.concentrations() - C) / np.abs(C))))
        if np.max(np.abs(mc)) > 0:
            worst_s = max(worst_s, float(np.max(np.abs(R.sources() - mc)) / np.max(np.abs(mc))))
        else:
            assert np.all(R.sources() == 0)
    assert worst_c <= 1e-4 and worst_s <= 5e-5, (worst_c, worst_s)
    c = o.counters()
    assert c["total_new"] > 50 and c["total_out"] > 100 and c["n_compactions"] >= 1, c
    assert np.max(np.abs(C - 2.0)) > 1e-2                 # feed and uptake did move the concentrations
    n_used = c["n_used"]
    a, b = R.get_particles(n_used), o.get_particles(n_used)
    assert np.array_equal(a["position"], b["position"]) and np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["age_div"], b["age_div"]) and np.array_equal(a["age_hyd"], b["age_hyd"])
    np.testing.assert_allclose(a["props"], b["props"], rtol=1e-6, atol=0)   # north_star: float properties to 1e-6 (bit-identical here)


def test_two_phase_coupled_loop_of_the_reference_simulation_unit(orc, synth):
    """Two-phase flow: the reference's SimulationUnit with a gas phase, gas and liquid feeds and a FixedKla mass-transfer
    model (setMtrModel, updateHydro with a gas state, update_feed for both phases, ode_step's two-phase branch: mass
    transfer, performStepGL gas then liquid, clearNegs) against orc.ode_step_gl + the oracle's cycle."""
    model, ns, n_comp, n, dt = "simple_acetate", 2, 30, 20_000, 0.5
    case = util.make_case(synth, model, n, n_comp, dt=dt, p_move=0.2, p_exit=0.0, outlet=False)
    fm = case["fm"]
    vl = np.ascontiguousarray(fm["volumes"], np.float64)
    fm_g = synth.make_flowmap(n_comp, dt, p_move=0.2, seed=8)
    rng = np.random.default_rng(3)
    vg = np.ascontiguousarray(0.03 * vl * (0.5 + rng.random(n_comp)))
    coo_g = (fm_g["coo"][0], fm_g["coo"][1], 0.01 * fm_g["coo"][2])
    q = 0.02 * vl[n_comp - 1] / dt
    qg = 0.1 * float(vg.min()) / dt
    lf = [dict(species=0, input_position=0, flow=q, concentration=5.0, output_position=n_comp - 1)]
    gf = [dict(species=1, input_position=0, flow=qg, concentration=0.28, output_position=n_comp - 1)]
    Cl = np.ascontiguousarray(np.stack([2.0 + rng.random(n_comp), 1e-3 * rng.random(n_comp)], axis=1).ravel())
    Cg = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.25 + 0.05 * rng.random(n_comp)], axis=1).ravel())
    kla_fixed = [0.0, 0.05]
    w = case["weight"] * 2e3
    R = refmod.RefSim(model, ns, n_comp, vl, Cl, lf, seed=11, gas=dict(volumes=vg, c0=Cg, feeds=gf, kla_fixed=kla_fixed))
    R.set_gas_hydro(vg, coo_g); R.update_hydro(fm)
    R.set_particles(case["props"], case["pos"], w)
    o = orc.OracleLoop(model, ns, n_comp, seed=11, n_threads=2)
    o.set_particles(case["props"], case["pos"]); o.set_weight(w)
    o.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    o.set_leaving_flows([(n_comp - 1, q, vl[n_comp - 1])])
    o.set_quirk_contrib_return(True)
    kla = np.tile(np.array(kla_fixed), n_comp); henry = np.array([0.0, 3.181e-2])
    ml, mg = Cl * np.repeat(vl, ns), Cg * np.repeat(vg, ns)
    mc = np.zeros(ns * n_comp)
    Cl0 = Cl.copy()
    for step in range(20):
        R.step(dt)
        sl = mc.copy(); sl[0] += q * 5.0
        kl = np.zeros(n_comp); kl[n_comp - 1] += q
        sg = np.zeros(ns * n_comp); sg[1] += qg * 0.28
        kg = np.zeros(n_comp); kg[n_comp - 1] += qg
        mtr = orc.ode_step_gl(Cl, ml, vl, kl, sl, fm["coo"], Cg, mg, vg, kg, sg, coo_g, kla, henry, dt)
        o.set_concentrations(Cl); o.cycle(dt)
        mc = o.get_sources().copy()
        util.assert_counters_equal(R.counters(), o.counters())
        g, m = R.gas()
        cl = R.concentrations()
        for s in range(ns):   # per species, relative to its largest value (dissolved oxygen passes near zero)
            assert np.max(np.abs(cl[s::ns] - Cl[s::ns])) <= 1e-4 * np.max(np.abs(Cl[s::ns])), (step, s)
        assert np.max(np.abs(g - Cg)) <= 1e-4 * np.max(np.abs(Cg)) and np.max(np.abs(m - mtr)) <= 1e-4 * np.max(np.abs(mtr)), step
        if step == 0:   # before any particle contribution exists the two are the same arithmetic
            assert np.array_equal(cl, Cl) and np.array_equal(g, Cg) and np.array_equal(m, mtr)
    assert o.counters()["total_out"] > 50 and np.any(mtr[1::2] != 0) and np.max(np.abs(Cl - Cl0)) > 1e-3
    n_used = o.counters()["n_used"]
    a, b = R.get_particles(n_used), o.get_particles(n_used)
    assert np.array_equal(a["position"], b["position"]) and np.array_equal(a["status"], b["status"])
    np.testing.assert_allclose(a["props"], b["props"], rtol=1e-5, atol=1e-30)
