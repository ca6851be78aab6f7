"""Every instantiation of the step kernel that the library ships is run through the parity cases.

Each model has its default block size and at most one alternative, selected by BMC_VARIANT=alt (read at bmc_create);
the eager-age kernel of a variant is a third instantiation (exercised by the cases with caller-supplied ages or a
changing time step).  `kernel_config()` proves which instantiation ran."""
import pytest

import test_ages_gpu as ta
import test_parity_gpu as tp
import test_reference_sources as tr

pytestmark = pytest.mark.gpu

# model -> (default block, alternative variant, its block): bmc_inst_*.cu
ALT = {"fixed_length": (1024, "v4b3", 768), "monod": (1024, "v4b3", 768), "simple_acetate": (768, "v4b4", 1024)}


@pytest.mark.parametrize("model", sorted(ALT))
def test_variant_selection(bmc, monkeypatch, model):
    ns = 2 if model == "simple_acetate" else 1
    assert bmc.ParticleLoop(model, ns, 8).kernel_config()["block"] == ALT[model][0]
    monkeypatch.setenv("BMC_VARIANT", ALT[model][1])
    k = bmc.ParticleLoop(model, ns, 8).kernel_config()
    assert k["block"] == ALT[model][2] and k["vec"] == 4
    # the 8-property wide model ships 1024 (default) and 768 ("v4b3") threads
    assert bmc.ParticleLoop("wide_udf", 4, 8, n_var_udf=8).kernel_config()["block"] == (768 if ALT[model][1] == "v4b3" else 1024)
    monkeypatch.setenv("BMC_VARIANT", "v9b9")   # unknown: ignored
    assert bmc.ParticleLoop(model, ns, 8).kernel_config()["block"] == ALT[model][0]


@pytest.mark.parametrize("model", ["fixed_length", "monod", "wide_udf"])
def test_alt_many_steps_division_exit_compaction(bmc, orc, synth, monkeypatch, model):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    tp.test_many_steps_division_exit_compaction(bmc, orc, synth, model)


def test_alt_simple_acetate(bmc, orc, synth, monkeypatch):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    tp.test_simple_acetate_deterministic_part(bmc, orc, synth)


def test_alt_large_compartment_tables(bmc, orc, synth, monkeypatch):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    tp.test_large_compartment_tables(bmc, orc, synth)


def test_alt_synchronised_population(bmc, orc, synth, monkeypatch):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    tp.test_synchronised_population_many_divisions_in_one_step(bmc, orc, synth)


@pytest.mark.parametrize("name", tr.NAMES)
def test_alt_reproduces_reference_fixture(bmc, monkeypatch, name):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    tr.test_cuda_reproduces_reference_fixture(bmc, name)


@pytest.mark.parametrize("model", ["fixed_length", "monod"])
def test_alt_stamped_ages(bmc, orc, synth, monkeypatch, model):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    ta.test_stamped_ages_many_steps(bmc, orc, synth, model)


def test_alt_eager_ages(bmc, orc, synth, monkeypatch):
    monkeypatch.setenv("BMC_VARIANT", "alt")
    ta.test_nonzero_initial_ages_use_the_eager_kernel(bmc, orc, synth)
    ta.test_time_step_change_switches_to_eager(bmc, orc, synth)
    ta.test_outlet_switched_on_mid_run(bmc, orc, synth)


def test_wide_models_default_instantiations(bmc, orc, synth):
    # WideUdf<16> (VEC 2), <32> and <64> (VEC 1) ship one block size each: many steps with division, exits, compaction
    import util
    for nv in (8, 16, 32, 64):
        case = util.make_case(synth, "wide_udf", 40_000, 100, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3, n_var_udf=nv)
        g, o = tp._pair(bmc, orc, case, dead_ratio=0.0005)
        assert g.kernel_config()["vec"] == {8: 4, 16: 2, 32: 1, 64: 1}[nv]
        util.load_case(g, case); util.load_case(o, case)
        for _ in range(2):
            sg = util.run_steps(g, case, 5, collect=True); so = util.run_steps(o, case, 5, collect=True)
            tp._compare_sources(sg, so)
            tp._compare(g, o)
        assert o.counters()["total_new"] > 0 and o.counters()["n_compactions"] >= 1


def _run_sources(bmc, synth, case, steps):
    import util
    g = bmc.ParticleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], dead_ratio=0.0005)
    util.load_case(g, case)
    src = util.run_steps(g, case, steps, collect=True)
    return g, src


@pytest.mark.parametrize("model", ["monod", "simple_acetate"])
def test_source_terms_are_bit_identical_across_runs_and_block_sizes(bmc, synth, monkeypatch, model):
    """The scatter adds 64-bit fixed-point integers (native shared-memory atomics, RED.ADD.64 flushes): the sums do not
    depend on the order of the additions.  From the second step on (the first step after a load has no scale yet and
    takes the fp64 path) the source terms of two runs are the same BITS — also when the block size, and with it the
    assignment of particles to blocks and warps, differs.  (The reference's ScatterView<float> is not reproducible.)"""
    import numpy as np
    import util
    case = util.make_case(synth, model, 150_000, 300, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    if model == "simple_acetate":
        case["props"][1, :] = 1.0   # no division: its division draws through libdevice, irrelevant here
    _, a = _run_sources(bmc, synth, case, 8)
    _, b = _run_sources(bmc, synth, case, 8)
    monkeypatch.setenv("BMC_VARIANT", "alt")
    g3, c = _run_sources(bmc, synth, case, 8)
    assert g3.kernel_config()["block"] == ALT[model][2]
    for k in range(1, 8):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    assert np.any(a[-1] != 0)


def test_fixed_point_scatter_keeps_small_contributions(bmc, orc, synth):
    """Substrate spanning seven decades across the compartments: the uptake of a starving compartment is 1e-7 of the
    richest one.  The fixed-point bins are scaled to the LARGEST contribution, so the absolute error per particle is
    bounded by 2^-35 of that (at this population far less): every bin agrees with the fp64 oracle to 1e-9 of the largest
    term, and bins down to 1e-6 of the largest still agree to 1e-6 of themselves."""
    import numpy as np
    import util
    n_comp = 200
    case = util.make_case(synth, "monod", 200_000, n_comp, dt=1.0, p_move=0.05, outlet=False)
    case["conc"] = 10.0 ** np.linspace(-9, -2, n_comp)          # k_s = 1e-3: mu from 1e-6 mu_max to 0.9 mu_max
    g = bmc.ParticleLoop("monod", 1, n_comp, seed=case["seed"]); o = orc.OracleLoop("monod", 1, n_comp, seed=case["seed"], n_threads=4)
    util.load_case(g, case); util.load_case(o, case)
    for step in range(6):
        g.set_concentrations(case["conc"]); o.set_concentrations(case["conc"])
        g.cycle(case["dt"]); o.cycle(case["dt"])
        sg, so = g.get_sources(), o.get_sources()
        big = np.max(np.abs(so))
        assert big > 0 and np.max(np.abs(sg - so)) <= 1e-9 * big
        if step >= 1:   # fixed-point path
            sel = np.abs(so) >= 1e-6 * big
            assert sel.sum() > n_comp // 2
            assert np.max(np.abs(sg[sel] - so[sel]) / np.abs(so[sel])) <= 1e-6
    tp._compare(g, o)


@pytest.mark.parametrize("model,n", [("fixed_length", 2_000_077), ("monod", 2_000_077), ("simple_acetate", 1_500_013), ("wide_udf", 400_019)])
def test_static_head_dynamic_tail_work_distribution(bmc, orc, synth, monkeypatch, model, n):
    """Work distribution of the particle pass (cycle_body): the first groups of a warp are its grid-stride groups, the
    last ones are drawn dynamically.  Small populations never reach the static part (fewer groups than warps), so this
    case is sized for two static rounds followed by the dynamic tail (BMC_DYN_MIN=1, BMC_DYN_SHIFT=2: a quarter of three
    groups per warp rounds to the minimum of one), on every kernel family — and must be bit-identical to the oracle like
    the fully dynamic scheme."""
    import util
    monkeypatch.setenv("BMC_DYN_MIN", "1")
    monkeypatch.setenv("BMC_DYN_SHIFT", "2")
    kw = dict(n_var_udf=32) if model == "wide_udf" else {}   # VEC = 1 kernel: the other loop text of cycle_body
    case = util.make_case(synth, model, n, 60, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3, **kw)
    if model == "simple_acetate":
        case["props"][1, :] = 1.0   # no division: its division draws through libdevice
    g, o = tp._pair(bmc, orc, case, dead_ratio=0.0005)
    k = g.kernel_config()
    groups = (n + 32 * k["vec"] - 1) // (32 * k["vec"])
    assert groups // (k["grid"] * k["block"] // 32) >= 3, "population too small to reach the static rounds on this device"
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 3, collect=True); so = util.run_steps(o, case, 3, collect=True)
    tp._compare_sources(sg, so)
    tp._compare(g, o)
    c = o.counters()
    assert c["total_out"] > 0 and (model == "simple_acetate" or c["total_new"] > 0)


@pytest.mark.parametrize("model,n", [("monod", 2_000_077), ("fixed_length", 1_200_031)])
def test_pass_without_prefetch_staging(bmc, orc, synth, monkeypatch, model, n):
    """The particle pass without the cp.async staging (BMC_PREFETCH=0; what a 10 000-compartment case runs, whose bins and
    table leave no room for it): tickets are resolved at the end of a group instead of in the middle.  Sized so that
    every block goes through the ring of chunk descriptors several times (a chunk size smaller than the number of warps
    once deadlocked exactly this path), bit-identical to the oracle."""
    import util
    monkeypatch.setenv("BMC_PREFETCH", "0")
    monkeypatch.setenv("BMC_DYN_SHIFT", "0")
    case = util.make_case(synth, model, n, 60, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    g, o = tp._pair(bmc, orc, case, dead_ratio=0.0005)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 3, collect=True); so = util.run_steps(o, case, 3, collect=True)
    tp._compare_sources(sg, so)
    tp._compare(g, o)
    assert o.counters()["total_out"] > 0 and o.counters()["total_new"] > 0
