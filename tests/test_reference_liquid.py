"""Liquid / gas scalar solver pinned to the REFERENCE'S OWN sources: apps/libs/simulation/src/implScalar.cpp
(ScalarSimulation::performStep, performStepGL, clearNegs, set_transition, set_mass) and src/hydro/mass_transfer.cpp
(MassTransferModel::gas_liquid_mass_transfer, update) and src/hydro/impl_mtr.cpp (the kla correlation of
Type::FlowmapTurbulence) are compiled where they lie over oracle/eigen_shim (Eigen is a system
package of the reference's build, absent from this image), oracle/rust_shim and oracle/kokkos_shim into
oracle/_ref/libbmc_ref_liquid.so; oracle/ref_liquid.cpp adds the per-step call sequence of the reference's main loop
(update_feed -> ode_step -> clearContribution).

CPU suite: the oracle's restatement (orc.ode_step / ode_step_gl) against that library, bit for bit, over long
trajectories with many feeds, sinks and Monte-Carlo source terms.  GPU suite: the CUDA path (bmc_liquid_step, one- and
two-phase) against the same library, bit for bit, through the C ABI."""
import numpy as np
import pytest

import ref as refmod

pytestmark = pytest.mark.skipif(not refmod.liquid_available(), reason="oracle/_ref/libbmc_ref_liquid.so not built (needs /root/reference)")


def _feeds(rng, ns, n_comp, n, scale):
    out = []
    for _ in range(n):
        fl = float(rng.random() * scale); ip = int(rng.integers(n_comp)); op = int(rng.integers(n_comp))
        for s in range(ns):
            out.append(dict(species=s, input_position=ip, flow=fl, concentration=float(rng.random() * 5), output_position=op, first_of_feed=int(s == 0)))
    return out


def _terms(feeds, ns, nc, mc=None):
    """`sources` and `sink` as the reference builds them: sources hold the Monte-Carlo contribution (synchro_sources),
    set_feed then adds flow * concentration feed by feed; set_sink adds the flow once per feed"""
    src = np.zeros(ns * nc) if mc is None else mc.copy()
    sink = np.zeros(nc)
    for f in feeds:
        src[f["species"] + ns * f["input_position"]] += f["flow"] * f["concentration"]
        if f.get("output_position") is not None and f.get("first_of_feed", 1):
            sink[f["output_position"]] += f["flow"]
    return src, sink


def _shuffled(coo, rng, duplicates=0):
    """the same matrix with its triplets in random order and some entries split in two (setFromTriplets sums them)"""
    r, c, v = (np.asarray(x).copy() for x in coo)
    if duplicates:
        k = rng.choice(v.size, size=min(duplicates, v.size), replace=False)
        r = np.concatenate([r, r[k]]); c = np.concatenate([c, c[k]]); v = np.concatenate([v, 0.25 * v[k]]); v[k] *= 0.75
    p = rng.permutation(v.size)
    return r[p], c[p], v[p]


@pytest.mark.parametrize("n_comp,ns,n_feeds,shuffle", [(1, 1, 1, False), (64, 1, 12, False), (300, 2, 12, True), (500, 2, 3, False)])
def test_oracle_liquid_step_equals_reference_solver(orc, synth, n_comp, ns, n_feeds, shuffle):
    fm = synth.make_flowmap(n_comp, 0.1, p_move=0.05)
    rng = np.random.default_rng(5)
    coo = _shuffled(fm["coo"], rng, duplicates=40) if shuffle else fm["coo"]
    C0 = rng.random(ns * n_comp) + 0.5
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    feeds = _feeds(rng, ns, n_comp, n_feeds, 1e-4)
    R = refmod.RefLiquid(ns, n_comp, vol); R.set_hydro(vol, coo); R.set_concentration(C0)
    C = C0.copy(); mass = C * np.repeat(vol, ns)
    for step in range(200):
        mc = -1e-6 * rng.random(ns * n_comp)                      # uptake by the particles of the previous cycle
        src, sink = _terms(feeds, ns, n_comp, mc)
        orc.ode_step(C, mass, vol, sink, src, coo, 0.1)
        R.step(0.1, mc, feeds)
        got = R.concentration()
        assert np.array_equal(got.view(np.uint64), C.view(np.uint64)), (step, np.max(np.abs(got - C)))
    assert np.max(np.abs(C - C0)) > 1e-4


@pytest.mark.parametrize("n_comp", [1, 64, 300])
def test_oracle_two_phase_step_equals_reference_solver(orc, synth, n_comp):
    ns, dt = 2, 0.05
    fm_l = synth.make_flowmap(n_comp, dt, p_move=0.05, seed=7); fm_g = synth.make_flowmap(n_comp, dt, p_move=0.2, seed=8)
    rng = np.random.default_rng(3)
    vl = np.ascontiguousarray(fm_l["volumes"], np.float64)
    vg = np.ascontiguousarray(0.03 * vl * (0.5 + rng.random(n_comp)), np.float64)
    Cl = np.ascontiguousarray(np.stack([2.0 + rng.random(n_comp), 1e-3 * rng.random(n_comp)], axis=1).ravel())
    Cg = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.25 + 0.05 * rng.random(n_comp)], axis=1).ravel())
    coo_g = (fm_g["coo"][0], fm_g["coo"][1], 0.01 * fm_g["coo"][2])
    kla = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.02 + 0.05 * rng.random(n_comp)], axis=1).ravel())   # a field, not a constant
    R = refmod.RefLiquid(ns, n_comp, vl); R.set_hydro(vl, fm_l["coo"]); R.set_concentration(Cl)
    R.enable_gas(vg, [0.0, 0.05]); R.set_hydro(vg, coo_g, gas=True); R.set_concentration(Cg, gas=True)
    henry = R.default_henry()
    assert henry.tolist() == [0.0, 3.181e-2]                      # MassTransferModel's constructor (mass_transfer.cpp:116-119)
    R.set_kla_henry(kla, henry)
    lf = _feeds(rng, ns, n_comp, 3, 2e-5)
    gf = [dict(species=1, input_position=0, flow=0.1 * float(vg.min()) / dt, concentration=0.28, output_position=n_comp - 1)]
    ml, mg = Cl * np.repeat(vl, ns), Cg * np.repeat(vg, ns)
    for step in range(120):
        mc = -1e-7 * rng.random(ns * n_comp)
        sl, kl_ = _terms(lf, ns, n_comp, mc); sg, kg = _terms(gf, ns, n_comp)
        mtr = orc.ode_step_gl(Cl, ml, vl, kl_, sl, fm_l["coo"], Cg, mg, vg, kg, sg, coo_g, kla, henry, dt)
        R.step(dt, mc, lf, gf)
        assert np.array_equal(R.mass_transfer().view(np.uint64), mtr.view(np.uint64)), step
        assert np.array_equal(R.concentration().view(np.uint64), Cl.view(np.uint64)), step
        assert np.array_equal(R.concentration(gas=True).view(np.uint64), Cg.view(np.uint64)), step
    assert np.any(mtr[1::2] != 0) and np.all(mtr[0::2] == 0)           # only oxygen (species 1) is transferred


def test_oracle_clear_negs_equals_reference_solver(orc):
    ns, dt = 2, 1.0
    vl, vg = np.array([0.02]), np.array([0.002])
    empty = (np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0))
    for c_o2, clipped in ((1e-7, True), (1e-5, False)):           # dt * kla > 1 overshoots: -5e-8 is clipped, -5e-6 is kept
        Cl, Cg = np.array([1.0, c_o2]), np.zeros(2)
        R = refmod.RefLiquid(ns, 1, vl); R.set_hydro(vl, empty); R.set_concentration(Cl)
        R.enable_gas(vg, [0.0, 1.5]); R.set_hydro(vg, empty, gas=True); R.set_concentration(Cg, gas=True)
        ml, mg = Cl * 0.02, Cg * 0.002
        orc.ode_step_gl(Cl, ml, vl, np.zeros(1), np.zeros(2), empty, Cg, mg, vg, np.zeros(1), np.zeros(2), empty, np.array([0.0, 1.5]), R.default_henry(), dt)
        R.step(dt)
        got = R.concentration()
        assert np.array_equal(got.view(np.uint64), Cl.view(np.uint64))
        assert (got[1] == 0.0) == clipped and (clipped or got[1] < 0.0)


def test_kla_turbulence_correlation_equals_reference(orc, synth):
    """Type::FlowmapTurbulence: kl = 0.3 (eps nu)^0.25 Sc^-0.5, a = 6 alpha / (db (1 - alpha)) (hydro/impl_mtr.cpp:22-149, the
    reference's own source over the Eigen stand-in) against the oracle's restatement and the host layer (cma.py) that
    feeds bmc_mass_transfer_set; then a two-phase trajectory driven by the reference's own kla field."""
    from _bmc_loader import load_pkg
    cma = __import__("importlib").import_module(load_pkg().__name__ + ".cma")
    ns, n_comp, dt = 2, 120, 0.05
    fm_l = synth.make_flowmap(n_comp, dt, p_move=0.05, seed=7); fm_g = synth.make_flowmap(n_comp, dt, p_move=0.2, seed=8)
    rng = np.random.default_rng(9)
    vl = np.ascontiguousarray(fm_l["volumes"], np.float64)
    vg = np.ascontiguousarray(0.03 * vl * (0.5 + rng.random(n_comp)), np.float64)
    eps = 0.2 + rng.random(n_comp)
    R = refmod.RefLiquid(ns, n_comp, vl); R.set_hydro(vl, fm_l["coo"])
    R.enable_gas_turbulence(vg)
    coo_g = (fm_g["coo"][0], fm_g["coo"][1], 0.01 * fm_g["coo"][2])
    R.set_hydro(vg, coo_g, gas=True)
    Cl = np.ascontiguousarray(np.stack([2.0 + rng.random(n_comp), 1e-3 * rng.random(n_comp)], axis=1).ravel())
    Cg = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.25 + 0.05 * rng.random(n_comp)], axis=1).ravel())
    R.set_concentration(Cl); R.set_concentration(Cg, gas=True)
    kla_ref = R.update_mass_transfer(vl, vg, eps)
    kla_orc = orc.kla_flowmap_turbulence(ns, eps, vl, vg)
    kla_host = cma.kla_flowmap_turbulence(ns, eps, vl, vg)
    assert np.all(kla_ref[0::2] == 0) and np.all(kla_ref[1::2] > 0)
    assert np.array_equal(kla_ref.view(np.uint64), kla_orc.view(np.uint64)), np.max(np.abs(kla_ref - kla_orc) / kla_ref.max())
    np.testing.assert_allclose(kla_host, kla_ref, rtol=1e-13, atol=0)     # numpy's pow against libm's
    henry = R.default_henry()
    ml, mg = Cl * np.repeat(vl, ns), Cg * np.repeat(vg, ns)
    z1, z2 = np.zeros(n_comp), np.zeros(ns * n_comp)
    for step in range(50):
        mtr = orc.ode_step_gl(Cl, ml, vl, z1, z2, fm_l["coo"], Cg, mg, vg, z1, z2.copy(), coo_g, kla_ref, henry, dt)
        R.step(dt)
        assert np.array_equal(R.mass_transfer().view(np.uint64), mtr.view(np.uint64)), step
        assert np.array_equal(R.concentration().view(np.uint64), Cl.view(np.uint64)), step
        assert np.array_equal(R.concentration(gas=True).view(np.uint64), Cg.view(np.uint64)), step


# ------------------------------------------------------------------------------------------------- CUDA path (C ABI)
@pytest.mark.gpu
@pytest.mark.parametrize("n_comp,ns,model,shuffle", [(1, 1, "monod", False), (64, 1, "monod", True), (500, 2, "simple_acetate", True)])
def test_cuda_liquid_step_equals_reference_solver(bmc, synth, n_comp, ns, model, shuffle):
    fm = synth.make_flowmap(n_comp, 0.1, p_move=0.05)
    rng = np.random.default_rng(11)
    coo = _shuffled(fm["coo"], rng, duplicates=25) if shuffle else fm["coo"]
    C0 = rng.random(ns * n_comp) + 0.5
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    feeds = _feeds(rng, ns, n_comp, 7, 1e-4)[:16]
    g = bmc.ParticleLoop(model, ns, n_comp)
    g.domain_update(fm["volumes"], fm["neighbors"] if n_comp > 1 else None, fm["out_flows"], fm["cdf"] if n_comp > 1 else None)
    g.liquid_set_transition(coo); g.set_concentrations(C0); g.liquid_set_feeds(feeds)
    R = refmod.RefLiquid(ns, n_comp, vol); R.set_hydro(vol, coo); R.set_concentration(C0)
    for step in range(60):
        g.liquid_step(0.1)
        R.step(0.1, None, feeds)
    got, want = g.get_concentrations(), R.concentration()
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), np.max(np.abs(got - want))
    assert np.max(np.abs(want - C0)) > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n_comp", [1, 64, 300])
def test_cuda_two_phase_step_equals_reference_solver(bmc, synth, n_comp):
    ns, dt = 2, 0.05
    fm_l = synth.make_flowmap(n_comp, dt, p_move=0.05, seed=7); fm_g = synth.make_flowmap(n_comp, dt, p_move=0.2, seed=8)
    rng = np.random.default_rng(3)
    vl = np.ascontiguousarray(fm_l["volumes"], np.float64)
    vg = np.ascontiguousarray(0.03 * vl * (0.5 + rng.random(n_comp)), np.float64)
    Cl0 = np.ascontiguousarray(np.stack([2.0 + rng.random(n_comp), 1e-3 * rng.random(n_comp)], axis=1).ravel())
    Cg0 = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.25 + 0.05 * rng.random(n_comp)], axis=1).ravel())
    coo_g = (fm_g["coo"][0], fm_g["coo"][1], 0.01 * fm_g["coo"][2])
    kla = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.02 + 0.05 * rng.random(n_comp)], axis=1).ravel())
    lf = _feeds(rng, ns, n_comp, 3, 2e-5)
    gf = [dict(species=1, input_position=0, flow=0.1 * float(vg.min()) / dt, concentration=0.28, output_position=n_comp - 1)]
    R = refmod.RefLiquid(ns, n_comp, vl); R.set_hydro(vl, fm_l["coo"]); R.set_concentration(Cl0)
    R.enable_gas(vg, [0.0, 0.05]); R.set_hydro(vg, coo_g, gas=True); R.set_concentration(Cg0, gas=True)
    henry = R.default_henry(); R.set_kla_henry(kla, henry)
    g = bmc.ParticleLoop("simple_acetate", ns, n_comp)
    g.domain_update(fm_l["volumes"], fm_l["neighbors"] if n_comp > 1 else None, fm_l["out_flows"], fm_l["cdf"] if n_comp > 1 else None)
    g.liquid_set_transition(fm_l["coo"]); g.set_concentrations(Cl0)
    g.gas_enable(vg, Cg0); g.gas_update_hydro(vg, coo_g); g.mass_transfer_set(kla, henry)
    g.liquid_set_feeds(lf); g.gas_set_feeds(gf)
    for step in range(60):
        g.liquid_step(dt)
        R.step(dt, None, lf, gf)
    assert np.array_equal(g.get_mass_transfer().view(np.uint64), R.mass_transfer().view(np.uint64))
    assert np.array_equal(g.get_concentrations().view(np.uint64), R.concentration().view(np.uint64))
    assert np.array_equal(g.get_gas_concentrations().view(np.uint64), R.concentration(gas=True).view(np.uint64))
