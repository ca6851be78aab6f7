"""Gas phase / gas-liquid mass transfer ("next" row 4 of SURVEY.md §8f): bmc_gas_* + the two-phase branch of
bmc_liquid_step against the oracle's restatement of SimulationUnit::ode_step (simulation.model.cpp:131-154,
implScalar.cpp:229-296, hydro/mass_transfer.cpp, hydro/impl_mtr.cpp).  Eigen is not in the image, so — like the
single-phase step — the oracle side is a restatement of the expressions, not the reference's own code."""
import importlib

import numpy as np
import pytest


def _cma(bmc):
    return importlib.import_module("biocma_mcst_b200.cma")


def test_kla_correlations_match_the_oracle(bmc, orc):
    cma = _cma(bmc)
    assert cma.c_kinematic_viscosity(20.0) == 1.0023e-06                 # water at 20 C, rounded to 1e-10 like the reference
    rng = np.random.default_rng(1)
    vl = 0.01 + 0.01 * rng.random(50); vg = 0.05 * vl * rng.random(50) + 1e-5; eps = 0.1 + rng.random(50)
    a = cma.kla_flowmap_turbulence(2, eps, vl, vg); b = orc.kla_flowmap_turbulence(2, eps, vl, vg)
    np.testing.assert_allclose(a, b, rtol=1e-13, atol=0)
    assert np.all(a[0::2] == 0) and np.all(a[1::2] > 0)                # only oxygen (species 1) is transferred
    assert cma.kla_fixed([0.0, 0.02], 3).tolist() == [0.0, 0.02] * 3 and cma.default_henry(2).tolist() == [0.0, 3.181e-2]


def _feed_terms(feeds, ns, nc):
    src = np.zeros(ns * nc); sink = np.zeros(nc)
    for f in feeds:
        src[f["species"] + ns * f["input_position"]] += f["flow"] * f["concentration"]
        if f.get("output_position") is not None and f.get("first_of_feed", 1):
            sink[f["output_position"]] += f["flow"]
    return src, sink


@pytest.mark.gpu
@pytest.mark.parametrize("n_comp,kind", [(1, "fixed"), (64, "fixed"), (300, "turbulence")])
def test_two_phase_step_bit_exact(bmc, orc, synth, n_comp, kind):
    cma = _cma(bmc)
    ns, dt = 2, 0.05
    fm_l = synth.make_flowmap(n_comp, dt, p_move=0.05, seed=7)
    fm_g = synth.make_flowmap(n_comp, dt, p_move=0.2, seed=8)            # the gas moves faster, on its own map
    rng = np.random.default_rng(3)
    vl = np.ascontiguousarray(fm_l["volumes"], np.float64)
    vg = np.ascontiguousarray(0.03 * vl * (0.5 + rng.random(n_comp)), np.float64)
    Cl0 = np.ascontiguousarray(np.stack([2.0 + rng.random(n_comp), 1e-3 * rng.random(n_comp)], axis=1).ravel())   # glucose, dissolved O2
    Cg0 = np.ascontiguousarray(np.stack([np.zeros(n_comp), 0.25 + 0.05 * rng.random(n_comp)], axis=1).ravel())   # O2 in the gas
    if kind == "fixed":
        kla = cma.kla_fixed([0.0, 0.05], n_comp)
    else:
        kla = cma.kla_flowmap_turbulence(ns, 0.2 + rng.random(n_comp), vl, vg)
    henry = cma.default_henry(ns)
    g = bmc.ParticleLoop("simple_acetate", ns, n_comp)
    g.domain_update(fm_l["volumes"], fm_l["neighbors"] if n_comp > 1 else None, fm_l["out_flows"], fm_l["cdf"] if n_comp > 1 else None)
    g.liquid_set_transition(fm_l["coo"])
    g.set_concentrations(Cl0)
    g.gas_enable(vg, Cg0)
    coo_g = (fm_g["coo"][0], fm_g["coo"][1], 0.01 * fm_g["coo"][2])   # gas flows sized for the (small) gas volumes: dt * F / V < 1
    g.gas_update_hydro(vg, coo_g)
    g.mass_transfer_set(kla, henry)
    lf = [dict(species=0, input_position=0, flow=2e-5, concentration=5.0, output_position=n_comp - 1)]
    gf = [dict(species=1, input_position=0, flow=0.1 * float(vg.min()) / dt, concentration=0.28, output_position=n_comp - 1)]   # sparger in, vent out
    g.liquid_set_feeds(lf); g.gas_set_feeds(gf)
    volx_l, volx_g = np.repeat(vl, ns), np.repeat(vg, ns)
    Cl, Cg = Cl0.copy(), Cg0.copy(); ml, mg = Cl * volx_l, Cg * volx_g
    for step in range(40):
        g.liquid_step(dt)
        sl, kl_ = _feed_terms(lf, ns, n_comp); sg, kg = _feed_terms(gf, ns, n_comp)
        mtr = orc.ode_step_gl(Cl, ml, vl, kl_, sl, fm_l["coo"], Cg, mg, vg, kg, sg, coo_g, kla, henry, dt)
    assert np.array_equal(g.get_concentrations().view(np.uint64), Cl.view(np.uint64)), np.max(np.abs(g.get_concentrations() - Cl))
    assert np.array_equal(g.get_gas_concentrations().view(np.uint64), Cg.view(np.uint64))
    assert np.array_equal(g.get_mass_transfer().view(np.uint64), mtr.view(np.uint64))
    assert np.all(mtr[0::2] == 0) and np.any(mtr[1::2] > 0)              # oxygen goes from the gas into the liquid
    assert np.all(np.isfinite(Cl)) and np.all(np.abs(Cg) < 10.0) and np.mean(Cl[1::2]) > np.mean(Cl0[1::2])   # stable; dissolved oxygen rose


@pytest.mark.gpu
def test_clear_negs_and_argument_checks(bmc, orc, synth):
    cma = _cma(bmc)
    ns, n_comp, dt = 2, 1, 1.0
    vl, vg = np.array([0.02]), np.array([0.002])
    g = bmc.ParticleLoop("simple_acetate", ns, n_comp)
    with pytest.raises(bmc.BmcError):
        g.mass_transfer_set(np.zeros(2))                                  # gas phase not enabled
    g.domain_update(vl, None, np.zeros(1), None)
    g.set_concentrations(np.array([1.0, 1e-7]))
    g.gas_enable(vg, np.zeros(2))
    g.mass_transfer_set(cma.kla_fixed([0.0, 1.5], 1), cma.default_henry(ns))   # dt * kla > 1: the explicit step overshoots
    g.liquid_step(dt)
    Cl, Cg = np.array([1.0, 1e-7]), np.zeros(2); ml, mg = Cl * 0.02, Cg * 0.002
    z1, z2 = np.zeros(1), np.zeros(2)
    empty = (np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0))
    orc.ode_step_gl(Cl, ml, vl, z1, z2, empty, Cg, mg, vg, z1, z2.copy(), empty, cma.kla_fixed([0.0, 1.5], 1), cma.default_henry(ns), dt)
    got = g.get_concentrations()
    assert got[1] == 0.0 and Cl[1] == 0.0                                  # -5e-8 clipped by clearNegs (|c| < 5e-7)
    assert np.array_equal(got.view(np.uint64), Cl.view(np.uint64))
    with pytest.raises(bmc.BmcError):
        g.gas_enable(np.array([0.0]))                                     # gas volumes must be positive
