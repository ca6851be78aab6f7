"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Deterministic mode (north_star): RNG-free model update/division, fixed flow map,
shared Philox uniforms for move/exit -> compartment indices, statuses and all
counters bit-exact; float properties and ages required to 1e-6 relative and in
fact compared BIT-EXACT (both sides are IEEE without contraction); source terms
to 1e-9 relative (fp64 accumulation, different summation order).
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

SRC_RTOL = 1e-9


def _pair(bmc, orc, case, **kw):
    g = bmc.ParticleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], n_var_udf=case["n_var_udf"], **kw)
    o = orc.OracleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], n_var_udf=case["n_var_udf"],
                       n_threads=4, **kw)
    return g, o


def _compare(g, o, exact=True):
    cg, co = g.counters(), o.counters()
    util.assert_counters_equal(cg, co)
    n = co["n_used"]
    util.assert_state_equal(g.get_particles(n), o.get_particles(n), n, exact_props=exact)
    assert np.array_equal(g.repartition(), o.repartition())


def _compare_sources(sg, so):
    for a, b in zip(sg, so):
        scale = np.max(np.abs(b)) + 1e-300
        assert np.max(np.abs(a - b)) <= SRC_RTOL * scale, np.max(np.abs(a - b)) / scale


@pytest.mark.parametrize("model", ["fixed_length", "monod", "simple_acetate", "wide_udf"])
def test_single_step_multi_compartment(bmc, orc, synth, model):
    case = util.make_case(synth, model, 50_000, 500, p_move=0.2, p_exit=0.05)
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 1, collect=True); so = util.run_steps(o, case, 1, collect=True)
    _compare_sources(sg, so)
    _compare(g, o)


@pytest.mark.parametrize("model", ["fixed_length", "monod", "wide_udf"])
def test_many_steps_division_exit_compaction(bmc, orc, synth, model):
    # dt large enough that cells divide, leave through the outlet, and the 1 % dead
    # threshold triggers compaction several times
    case = util.make_case(synth, model, 120_000, 500, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    g, o = _pair(bmc, orc, case, dead_ratio=0.0005)
    util.load_case(g, case); util.load_case(o, case)
    for blk in range(4):
        sg = util.run_steps(g, case, 5, collect=True); so = util.run_steps(o, case, 5, collect=True)
        _compare_sources(sg, so)
        _compare(g, o)
    c = o.counters()
    assert c["total_new"] > 0 and c["total_out"] > 0 and c["n_compactions"] >= 2, c


def test_simple_acetate_deterministic_part(bmc, orc, synth):
    # division of simple_acetate draws random numbers (stochastic mode); with no
    # division in range the whole step is deterministic and must be bit-exact
    case = util.make_case(synth, "simple_acetate", 80_000, 500, dt=0.5, p_move=0.2, p_exit=0.2)
    case["props"][1, :] = 1.0  # l_max far away: no division
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 12, collect=True); so = util.run_steps(o, case, 12, collect=True)
    _compare_sources(sg, so)
    _compare(g, o)
    assert o.counters()["total_new"] == 0


@pytest.mark.parametrize("model", ["fixed_length", "monod"])
def test_zero_d_batch_and_chemostat(bmc, orc, synth, model):
    # BASELINE configs[0]: 0D single compartment, 1e5 particles
    for outlet in (False, True):
        case = util.make_case(synth, model, 100_000, 1, dt=5.0, near_division=0.7, outlet=outlet, p_exit=0.02)
        g, o = _pair(bmc, orc, case)
        util.load_case(g, case); util.load_case(o, case)
        sg = util.run_steps(g, case, 10, collect=True); so = util.run_steps(o, case, 10, collect=True)
        _compare_sources(sg, so)
        _compare(g, o)
        if outlet:
            assert o.counters()["total_out"] > 0
        assert o.counters()["events"]["Move"] == 0


@pytest.mark.parametrize("n", [1, 3, 31, 1023, 1024, 1025, 4097])
def test_ragged_sizes(bmc, orc, synth, n):
    # the reference refuses N <= 1024 (kernels.hpp:163-167, Q8); we accept any N
    case = util.make_case(synth, "monod", n, 16, dt=30.0, near_division=0.9, p_move=0.5, p_exit=0.3)
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 6, collect=True); so = util.run_steps(o, case, 6, collect=True)
    _compare_sources(sg, so)
    _compare(g, o)


def test_initial_inactive_particles_and_forced_compaction(bmc, orc, synth):
    case = util.make_case(synth, "fixed_length", 30_000, 64, dt=1.0, p_move=0.3, outlet=False)
    rng = np.random.default_rng(7)
    status = np.where(rng.random(case["n"]) < 0.3, 2, 0).astype(np.uint8)
    status[-5:] = 2
    g, o = _pair(bmc, orc, case, dead_ratio=0.9)
    util.load_case(g, case, status); util.load_case(o, case, status)
    util.run_steps(g, case, 2); util.run_steps(o, case, 2)
    _compare(g, o)                      # no compaction yet (threshold 90 %)
    assert o.counters()["n_compactions"] == 0
    g.compact(); o.compact()            # force_remove_dead
    _compare(g, o)
    assert g.counters()["n_inactive"] == 0
    util.run_steps(g, case, 2); util.run_steps(o, case, 2)
    _compare(g, o)


def test_all_particles_exit(bmc, orc, synth):
    case = util.make_case(synth, "fixed_length", 5_000, 1, dt=1.0, outlet=True, p_exit=1e9)
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 2); util.run_steps(o, case, 2)
    cg, co = g.counters(), o.counters()
    util.assert_counters_equal(cg, co)


def test_division_buffer_overflow_counts(bmc, orc, synth):
    # every cell divides in the first step; buffer_ratio 0.1 cannot hold them.
    # Which mothers overflow is order dependent (also in the reference), the counts are not.
    case = util.make_case(synth, "fixed_length", 20_000, 8, dt=1.0, outlet=False)
    case["props"][0, :] = 2.5e-6
    kw = dict(buffer_ratio=0.1, allocation_factor=1.5)
    g, o = _pair(bmc, orc, case, **kw)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 1); util.run_steps(o, case, 1)
    cg, co = g.counters(), o.counters()
    assert cg["buffer_capacity"] >= co["buffer_capacity"]  # device rounds capacity up to whole tiles
    assert cg["events"]["NewParticle"] == co["events"]["NewParticle"] == 20_000
    assert cg["events"]["Overflow"] == 20_000 - cg["total_new"]
    assert cg["last_waiting_allocation"] == cg["events"]["Overflow"] > 0
    assert cg["n_used"] == 20_000 + cg["total_new"]
    # overflowed mothers keep l >= l_max and retry: after enough steps everybody has divided once
    for _ in range(12):
        util.run_steps(g, case, 1)
    assert g.counters()["total_new"] == 20_000 or g.counters()["n_used"] >= 40_000


def test_philox_stream_identity(bmc, orc, synth):
    # different seed / rank -> different trajectories; same seed -> identical reruns
    case = util.make_case(synth, "monod", 20_000, 100, p_move=0.3, p_exit=0.2)
    outs = []
    for seed, rank in ((1, 0), (1, 0), (2, 0), (1, 1)):
        g = bmc.ParticleLoop("monod", 1, 100, seed=seed, rank=rank)
        util.load_case(g, case)
        util.run_steps(g, case, 3)
        outs.append(g.get_particles()["position"].copy())
        o = orc.OracleLoop("monod", 1, 100, seed=seed, rank=rank)
        util.load_case(o, case); util.run_steps(o, case, 3)
        assert np.array_equal(outs[-1], o.get_particles()["position"])
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])
    assert not np.array_equal(outs[0], outs[3])


def test_particle_balance_identity(bmc, synth):
    # apps/core/src/post_process.cpp:92-117: sum(repartition) == new - removed + N0
    case = util.make_case(synth, "monod", 200_000, 500, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    g = bmc.ParticleLoop("monod", 1, 500)
    util.load_case(g, case)
    util.run_steps(g, case, 15)
    c = g.counters()
    assert int(g.repartition().sum()) == c["total_new"] - c["total_out"] + case["n"]
    assert c["n_used"] - c["n_inactive"] == int(g.repartition().sum())


def test_device_init_matches_oracle_deterministic_model(bmc, orc, synth):
    # mc_init_first: fixed_length is configurable (lengths given) -> positions (integer) bit-exact, mass to 1e-12
    n, nc = 50_000, 500
    rng = np.random.default_rng(3)
    linit = (1e-6 + 1e-6 * rng.random(n)).astype(np.float32)
    g = bmc.ParticleLoop("fixed_length", 1, nc, seed=77)
    o = orc.OracleLoop("fixed_length", 1, nc, seed=77)
    mg = g.init_particles(n, True, linit); mo = o.init_particles(n, True, linit)
    assert abs(mg - mo) <= 1e-12 * abs(mo)
    util.assert_state_equal(g.get_particles(n), o.get_particles(n), n)
    assert len(np.unique(g.get_particles(n)["position"])) == nc


# ---------------------------------------------------------------------------------------------
# Stochastic mode (north_star): agreement in distribution — two-sample KS on property histograms,
# occupancy and event counts within a stated tolerance.  simple_acetate's division draws
# LogNormal/TruncatedNormal variates through exp/log/erfc, which differ in the last bit between
# CUDA and glibc, so trajectories are not bit-comparable once cells have divided.
# ---------------------------------------------------------------------------------------------
KS_P_MIN = 1e-3        # reject only on strong evidence
COUNT_RTOL = 0.01      # event counts / occupancy within 1 %


def test_simple_acetate_division_in_distribution(bmc, orc, synth):
    from scipy.stats import ks_2samp
    case = util.make_case(synth, "simple_acetate", 200_000, 100, dt=30.0, near_division=0.6, p_move=0.2, p_exit=0.2)
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 25); util.run_steps(o, case, 25)
    cg, co = g.counters(), o.counters()
    assert co["total_new"] > 20_000
    for k in ("total_new", "total_out", "n_used"):
        assert abs(cg[k] - co[k]) <= COUNT_RTOL * co[k] + 30, (k, cg[k], co[k])
    assert abs(cg["events"]["Move"] - co["events"]["Move"]) <= COUNT_RTOL * co["events"]["Move"]
    pg, po = g.get_particles(), o.get_particles()
    ig, io = pg["status"] == 0, po["status"] == 0
    for col, name in ((0, "length"), (1, "l_max"), (2, "a_p"), (4, "a_e")):
        r = ks_2samp(pg["props"][col][ig].astype(np.float64), po["props"][col][io].astype(np.float64))
        assert r.pvalue > KS_P_MIN, (name, r)
    occ_g, occ_o = g.repartition().astype(np.float64), o.repartition().astype(np.float64)
    assert np.max(np.abs(occ_g - occ_o)) <= 6 * np.sqrt(occ_o.max())  # compartment occupancy
    sg, so = g.get_sources(), o.get_sources()
    assert abs(sg.sum() - so.sum()) <= 0.01 * abs(so.sum())            # total uptake within 1 %


def test_monod_device_init_in_distribution(bmc, orc):
    from scipy.stats import ks_2samp, chisquare
    n, nc = 400_000, 200
    g = bmc.ParticleLoop("monod", 1, nc, seed=5)
    o = orc.OracleLoop("monod", 1, nc, seed=5)
    mg = g.init_particles(n, True); mo = o.init_particles(n, True)
    assert abs(mg - mo) <= 1e-4 * mo                     # total mass (TruncatedNormal lengths)
    pg, po = g.get_particles(n), o.get_particles(n)
    assert ks_2samp(pg["props"][0].astype(np.float64), po["props"][0].astype(np.float64)).pvalue > KS_P_MIN
    lo, hi = np.float32(1e-6), np.float32(2e-6)
    assert pg["props"][0].min() > lo * 0.98 and pg["props"][0].max() < hi * 1.02
    assert np.all(pg["props"][1] == np.float32(2e-6)) and np.all(pg["props"][2] == np.float32(0.77 / 3600.0))
    assert np.array_equal(pg["position"], po["position"])  # urand64 from the same Philox words: exact
    assert chisquare(np.bincount(pg["position"].astype(np.int64), minlength=nc)).pvalue > 1e-4


def test_large_compartment_tables(bmc, orc, synth):
    # 4 species x 4000 compartments x 8 B = 128 KB of source bins: does not fit the shared-memory
    # budget -> the scatter falls back to L2 atomics; 10k x 1 species (80 KB) still uses shared bins
    for model, n_comp in (("wide_udf", 4000), ("monod", 10_000)):
        case = util.make_case(synth, model, 60_000, n_comp, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
        g, o = _pair(bmc, orc, case, dead_ratio=0.0005)
        util.load_case(g, case); util.load_case(o, case)
        sg = util.run_steps(g, case, 6, collect=True); so = util.run_steps(o, case, 6, collect=True)
        _compare_sources(sg, so)
        _compare(g, o)
        assert o.counters()["total_new"] > 0


def test_capacity_growth_is_transparent(bmc, orc, synth):
    # The population grows 5x with the reference's DEFAULT runtime parameters (allocation factor 1.5, buffer ratio 0.6)
    # and NO host synchronisation between the steps: the host follows the device through the pinned mirror, enlarges
    # the arrays ahead of the population, and the device follows the reference's logical extents (capacity(), buffer
    # extent: ParticlesContainer::_resize / __allocate_buffer__, particles_container.hpp:601-685) exactly.
    case = util.make_case(synth, "fixed_length", 30_000, 16, dt=300.0, near_division=0.5, p_move=0.3, outlet=False)
    g, o = _pair(bmc, orc, case, allocation_factor=1.5, buffer_ratio=0.6)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 40); util.run_steps(o, case, 40)
    cg, co = g.counters(), o.counters()
    assert co["n_used"] > 4 * case["n"] and cg["n_reallocations"] >= 2 and cg["events"]["Overflow"] == co["events"]["Overflow"]
    assert cg["capacity"] == co["capacity"] and cg["buffer_capacity"] == co["buffer_capacity"]   # the reference's extents, exactly
    assert cg["physical_capacity"] >= cg["capacity"]
    util.assert_counters_equal(cg, co)
    util.assert_state_equal(g.get_particles(co["n_used"]), o.get_particles(co["n_used"]), co["n_used"])


def _burst_case(synth, n=200_000):
    # quiet for three steps, then EVERY cell divides in the same step (a synchronised culture)
    case = util.make_case(synth, "fixed_length", n, 16, dt=100.0, p_move=0.3, outlet=False)
    case["props"][0][:] = np.float32(1.9995e-6)
    return case


def test_capacity_exact_mode_survives_a_synchronised_burst(bmc, orc, synth, monkeypatch):
    # BMC_CAPACITY_MODE=exact: lock-step with the device and worst-case room.  The reference's births are limited by the
    # division buffer only (0.6 * 1.5 * n here): 180 000 of the 200 000 cells divide at once, 20 000 are counted as
    # Overflow and divide one step later.  WHICH mothers overflow depends on the execution order in the reference too,
    # so the comparison is on the tallies and the extents, not on the particle rows.
    monkeypatch.setenv("BMC_CAPACITY_MODE", "exact")
    case = _burst_case(synth)
    g, o = _pair(bmc, orc, case, allocation_factor=1.5, buffer_ratio=0.6)
    util.load_case(g, case); util.load_case(o, case)
    for _ in range(6):
        util.run_steps(g, case, 2); util.run_steps(o, case, 2)
        cg, co = g.counters(), o.counters()
        for k in ("n_used", "capacity", "buffer_capacity", "total_new", "last_waiting_allocation"):
            assert cg[k] == co[k], (k, cg[k], co[k])
        for k in ("NewParticle", "Overflow"):
            assert cg["events"][k] == co["events"][k], (k, cg["events"], co["events"])
    assert co["events"]["Overflow"] == 20_000 and co["n_used"] == 2 * case["n"]


def test_burst_without_overflow_is_bit_exact_in_default_mode(bmc, orc, synth):
    # the same synchronised burst with a buffer that holds it (allocation factor 4): every cell divides in the very
    # first step, i.e. on an idle stream, where the room is provably sufficient -> bit-exact, no error, default mode
    case = _burst_case(synth)
    g, o = _pair(bmc, orc, case, allocation_factor=4.0, buffer_ratio=0.6)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 8); util.run_steps(o, case, 8)
    _compare(g, o)
    assert o.counters()["n_used"] == 2 * case["n"] and o.counters()["events"]["Overflow"] == 0


def test_capacity_exhaustion_fails_loudly(bmc, synth):
    # default mode: three steps without substrate (nothing grows, nothing divides) let the host run ahead; then the
    # substrate arrives and, with a 30-minute time step, EVERY cell divides in EVERY step — the population doubles per
    # step, which no history announced.  The first doubling fits (the arrays always hold the worst case of one step),
    # the second one, already enqueued, does not: the device refuses divisions for lack of PHYSICAL room, which the
    # reference would not have done.  That must surface as an error, never as a silent Overflow.
    n = 200_000
    case = util.make_case(synth, "fixed_length", n, 16, dt=1800.0, p_move=0.3, outlet=False)
    case["props"][0][:] = np.float32(1.9995e-6)
    g = bmc.ParticleLoop("fixed_length", 1, 16, allocation_factor=1.5, buffer_ratio=0.6)
    util.load_case(g, case)
    with pytest.raises(bmc.BmcError, match="capacity exhausted"):
        for step in range(12):
            g.set_concentrations(np.zeros(16) if step < 3 else np.full(16, 5.0))
            g.cycle(case["dt"])
        g.counters()
    # the same run in exact mode goes through (lock-step, reallocation before every step that needs it)


def test_synchronised_population_many_divisions_in_one_step(bmc, orc, synth):
    """> 2048 divisions in one step: newborn placement goes through the per-tile prefix path (mask popcounts +
    block prefix behind a grid barrier) instead of the direct ranking used for small batches; both must give
    ascending-mother order.  Then more steps with exits and compaction on the grown population."""
    n = 70_000
    case = util.make_case(synth, "fixed_length", n, 64, dt=600.0, p_move=0.3, p_exit=0.2)
    rng = np.random.default_rng(9)
    case["props"][0] = (1.9e-6 + 0.1e-6 * rng.random(n)).astype(np.float32)   # all close to l_max: about half divide at once
    g, o = _pair(bmc, orc, case, dead_ratio=0.002, allocation_factor=4.0)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 1, collect=True); so = util.run_steps(o, case, 1, collect=True)
    _compare_sources(sg, so)
    _compare(g, o)
    assert g.counters()["total_new"] > 4096
    # the whole population divides again two steps later (139k -> 278k); stop before the third doubling, which
    # exceeds the division buffer of the fixed 4x allocation (an overflow, handled differently by design)
    sg = util.run_steps(g, case, 4, collect=True); so = util.run_steps(o, case, 4, collect=True)
    _compare_sources(sg, so)
    _compare(g, o)
    assert g.counters()["total_new"] > 200_000 and g.counters()["n_compactions"] >= 3


def test_small_and_large_insert_paths_agree_on_order(bmc, synth):
    """the same population stepped with 1500 and with 5000 simultaneous divisions keeps newborns in ascending-mother order"""
    for n_div in (1500, 5000):
        n = 40_000
        case = util.make_case(synth, "fixed_length", n, 8, dt=1.0, outlet=False)
        case["props"][0][:] = 1.0e-6
        idx = np.sort(np.random.default_rng(n_div).choice(n, n_div, replace=False))
        case["props"][0][idx] = 2.5e-6                                         # exactly these divide in step 1
        g = bmc.ParticleLoop("fixed_length", 1, 8, allocation_factor=3.0)
        util.load_case(g, case)
        g.cycle(1.0)
        c = g.counters()
        assert c["total_new"] == n_div and c["n_used"] == n + n_div
        p = g.get_particles()
        # newborn k (slot n+k) is the daughter of the k-th dividing mother: same halved length, same compartment
        assert np.array_equal(p["props"][0][n:], p["props"][0][idx])
        assert np.array_equal(p["position"][n:], case["pos"][idx]) or case["n_comp"] > 1
