"""Step-stamped ages (bmc_kernels.cuh): while d_t and the outlets are constant and every age started at
zero, the age columns hold step stamps and are turned into floats only when read; otherwise the eager
kernel adds d_t every step like the reference (model_kernel.hpp:191, move_kernel.hpp:596).  Both forms,
and the switch between them, must give the oracle's eagerly accumulated ages BIT-exactly."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _pair(bmc, orc, case, **kw):
    g = bmc.ParticleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], **kw)
    o = orc.OracleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], n_threads=4, **kw)
    return g, o


def _same(g, o):
    cg, co = g.counters(), o.counters()
    util.assert_counters_equal(cg, co)
    n = co["n_used"]
    util.assert_state_equal(g.get_particles(n), o.get_particles(n), n)


def _load(loop, case, status=None, age_hyd=None, age_div=None):
    util.load_case(loop, case, status)
    if age_hyd is not None or age_div is not None or status is not None:
        loop.set_particles(case["props"], case["pos"], status, age_hyd, age_div)


@pytest.mark.parametrize("model", ["fixed_length", "monod"])
def test_stamped_ages_many_steps(bmc, orc, synth, model):
    # 60 steps with division, outlet exits and compactions: ages of mothers, newborns, exited and moved particles
    case = util.make_case(synth, model, 40_000, 50, dt=15.0, near_division=0.85, p_move=0.3, p_exit=0.2)
    # allocation_factor 4: room for the whole growth, so no division waits for a buffer slot (the lazily
    # grown device container and the reference's resize-on-merge differ only in WHEN an overflow happens)
    g, o = _pair(bmc, orc, case, dead_ratio=0.002, allocation_factor=4.0)
    _load(g, case); _load(o, case)
    for _ in range(6):
        util.run_steps(g, case, 10); util.run_steps(o, case, 10)
        _same(g, o)
    c = g.counters()
    assert c["total_new"] > 0 and c["total_out"] > 0 and c["n_compactions"] >= 1
    ages = g.get_particles()["age_div"]
    assert np.unique(ages).size > 10     # a real distribution of ages, not all equal


def test_nonzero_initial_ages_use_the_eager_kernel(bmc, orc, synth):
    case = util.make_case(synth, "monod", 30_000, 40, dt=10.0, near_division=0.8, p_move=0.3, p_exit=0.2)
    rng = np.random.default_rng(3)
    ah = (100.0 * rng.random(case["n"])).astype(np.float32); ad = (50.0 * rng.random(case["n"])).astype(np.float32)
    g, o = _pair(bmc, orc, case, dead_ratio=0.002)
    _load(g, case, None, ah, ad); _load(o, case, None, ah, ad)
    for _ in range(3):
        util.run_steps(g, case, 5); util.run_steps(o, case, 5)
        _same(g, o)


def test_initial_inactive_particles_have_frozen_ages(bmc, orc, synth):
    case = util.make_case(synth, "fixed_length", 20_000, 16, dt=5.0, p_move=0.3, p_exit=0.1)
    status = np.zeros(case["n"], np.uint8); status[::7] = 2
    g, o = _pair(bmc, orc, case, dead_ratio=0.9)     # no compaction: the inactive slots stay where they are
    _load(g, case, status); _load(o, case, status)
    util.run_steps(g, case, 8); util.run_steps(o, case, 8)
    _same(g, o)
    p = g.get_particles()
    assert np.all(p["age_div"][::7][p["status"][::7] == 2] == 0.0)


def test_time_step_change_switches_to_eager(bmc, orc, synth):
    case = util.make_case(synth, "monod", 30_000, 40, dt=10.0, near_division=0.8, p_move=0.3, p_exit=0.2)
    g, o = _pair(bmc, orc, case, dead_ratio=0.002)
    _load(g, case); _load(o, case)
    util.run_steps(g, case, 7); util.run_steps(o, case, 7)
    _same(g, o)
    util.run_steps(g, case, 6, dt=3.5); util.run_steps(o, case, 6, dt=3.5)      # stamps -> floats in place
    _same(g, o)
    util.run_steps(g, case, 4, dt=10.0); util.run_steps(o, case, 4, dt=10.0)
    _same(g, o)


def test_outlet_switched_on_mid_run(bmc, orc, synth):
    # age_hyd only advances while an outlet exists (cycle_move_leave runs only if enable_leave, kernels.hpp:142-157).
    # The hydraulic age has its own clock (steps WITH an outlet), so a fed-batch run that opens and closes its outlet
    # stays on the stamped-age kernel — and stays bit-identical to the eager accumulation of the oracle.
    case = util.make_case(synth, "fixed_length", 20_000, 16, dt=300.0, near_division=0.5, p_move=0.3, p_exit=0.1)
    flows = case["flows"]
    g, o = _pair(bmc, orc, case)
    _load(g, case); _load(o, case)
    g.set_leaving_flows([]); o.set_leaving_flows([])
    util.run_steps(g, case, 5); util.run_steps(o, case, 5)
    _same(g, o)
    assert np.all(g.get_particles()["age_hyd"] == 0.0)
    for k, fl in enumerate((flows, [], flows, [], flows)):   # open, close, open, ... with divisions and exits in between
        g.set_leaving_flows(fl); o.set_leaving_flows(fl)
        util.run_steps(g, case, 3 + k); util.run_steps(o, case, 3 + k)
        _same(g, o)
    assert np.any(g.get_particles()["age_hyd"] > 0.0)
    c = o.counters()
    assert c["total_new"] > 0 and c["total_out"] > 0
    assert g.kernel_config()["stamped_ages"], "an outlet toggle must not drop to the eager-age kernel"


def test_eager_forced_by_environment_matches(bmc, orc, synth, monkeypatch):
    monkeypatch.setenv("BMC_EAGER_AGES", "1")
    case = util.make_case(synth, "fixed_length", 20_000, 16, dt=15.0, near_division=0.85, p_move=0.3, p_exit=0.2)
    g, o = _pair(bmc, orc, case, dead_ratio=0.002)
    _load(g, case); _load(o, case)
    util.run_steps(g, case, 12); util.run_steps(o, case, 12)
    _same(g, o)
