"""Checkpoint / resume (SerDe::save_simulation / load_simulation, apps/core/src/serde.cpp:64-219).

The property the reference's serde exists for: a run resumed from a checkpoint continues exactly like the
uninterrupted one.  Checked bit-exactly — particle state, both ages, counters, tallies, sources — against an
uninterrupted CUDA run AND against the oracle, for the stamped and the eager age representation.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _mk(bmc, case, **kw):
    return bmc.ParticleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], dead_ratio=0.0005, **kw)


@pytest.mark.parametrize("model,n,n_comp,kw", [
    ("monod", 60_000, 500, dict(dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)),
    ("fixed_length", 20_000, 1, dict(dt=20.0, near_division=0.7, p_exit=0.05)),
])
def test_resume_is_bit_identical(bmc, orc, synth, model, n, n_comp, kw, monkeypatch):
    case = util.make_case(synth, model, n, n_comp, **kw)
    a = _mk(bmc, case)
    o = orc.OracleLoop(case["model"], case["n_species"], case["n_comp"], seed=case["seed"], n_threads=4, dead_ratio=0.0005)
    util.load_case(a, case); util.load_case(o, case)
    util.run_steps(a, case, 11); util.run_steps(o, case, 11)
    blob = a.checkpoint()
    ca = a.counters()
    assert ca["total_new"] > 0 and ca["total_out"] > 0

    # a fresh context: the flow map comes from the case (as in the reference), the unit from the checkpoint
    b = _mk(bmc, case)
    fm = case["fm"]
    if n_comp > 1:
        b.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    else:
        b.domain_update(fm["volumes"], None, fm["out_flows"], None)
    b.set_leaving_flows(case["flows"])
    b.restore(blob)
    cb = b.counters()
    util.assert_counters_equal(ca, cb)
    assert cb["step"] == 11
    util.assert_state_equal(a.get_particles(), b.get_particles(), ca["n_used"])
    assert np.array_equal(a.get_sources(), b.get_sources())

    def more(loop, first):
        out = []
        for s in range(first, first + 14):
            loop.set_concentrations(util.conc_at(case, s))
            loop.cycle(case["dt"])
            out.append(loop.get_sources().copy())
        return out
    sa, sb, so = more(a, 11), more(b, 11), more(o, 11)
    ca, cb, co = a.counters(), b.counters(), o.counters()
    util.assert_counters_equal(ca, cb); util.assert_counters_equal(cb, co)
    nu = ca["n_used"]
    util.assert_state_equal(a.get_particles(), b.get_particles(), nu)
    util.assert_state_equal(b.get_particles(), o.get_particles(nu), nu)
    for x, y, z in zip(sa, sb, so):
        assert np.max(np.abs(x - y)) <= 1e-12 * (np.max(np.abs(x)) + 1e-300)
        assert np.max(np.abs(y - z)) <= 1e-9 * (np.max(np.abs(z)) + 1e-300)
    assert ca["n_compactions"] >= 2


def test_resume_with_eager_ages(bmc, orc, synth):
    # a change of d_t switches the age columns to floats; the checkpoint carries that representation
    case = util.make_case(synth, "monod", 30_000, 200, dt=10.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    a = _mk(bmc, case)
    o = orc.OracleLoop("monod", 1, 200, seed=case["seed"], dead_ratio=0.0005)
    util.load_case(a, case); util.load_case(o, case)
    for dt in (10.0, 10.0, 5.0, 5.0):
        a.cycle(dt); o.cycle(dt)
    blob = a.checkpoint()
    b = _mk(bmc, case)
    util.load_case(b, case)   # domain + flows (+ particles that the checkpoint replaces)
    b.restore(blob)
    for dt in (5.0, 7.0, 7.0):
        a.cycle(dt); b.cycle(dt); o.cycle(dt)
    nu = o.counters()["n_used"]
    util.assert_counters_equal(a.counters(), b.counters()); util.assert_counters_equal(b.counters(), o.counters())
    util.assert_state_equal(a.get_particles(), b.get_particles(), nu)
    util.assert_state_equal(b.get_particles(), o.get_particles(nu), nu)


def test_checkpoint_rejects_other_model_and_garbage(bmc, synth):
    case = util.make_case(synth, "monod", 5_000, 20)
    a = _mk(bmc, case)
    util.load_case(a, case); a.cycle(case["dt"])
    blob = a.checkpoint()
    other = bmc.ParticleLoop("fixed_length", 1, 20)
    with pytest.raises(RuntimeError):
        other.restore(blob)
    with pytest.raises(RuntimeError):
        a.restore(b"not a checkpoint" * 50)
    with pytest.raises(RuntimeError):
        a.restore(blob[: len(blob) // 2])
    a.restore(blob)  # still usable
    a.cycle(case["dt"])


def test_force_remove_dead_keeps_last_sources(bmc, synth):
    # ParticlesContainer::force_remove_dead touches the container only: the source terms of the last cycle
    # (still to be consumed by the next ODE step, host_specific.cpp:257-313) must survive it
    case = util.make_case(synth, "monod", 20_000, 50, dt=20.0, near_division=0.5, p_move=0.2, p_exit=0.3)
    g = _mk(bmc, case)
    util.load_case(g, case)
    util.run_steps(g, case, 3)
    before = g.get_sources().copy()
    assert np.any(before != 0)
    g.compact()
    assert np.array_equal(g.get_sources(), before)
