"""Shared helpers: build one synthetic case and drive a loop (oracle or CUDA) through it."""
import numpy as np

N_SPECIES = {"fixed_length": 1, "monod": 1, "simple_acetate": 2, "wide_udf": 4, "udf_model": 1}


def make_case(synth, model, n, n_comp, *, dt=0.1, seed=2024, near_division=0.0, outlet=True, p_move=0.01,
              p_exit=1e-3, conc=5.0, n_var_udf=32, x0=0.5):
    fm = synth.make_flowmap(n_comp, dt, p_move=p_move, seed=seed)
    props, pos = synth.make_population(model, n, n_comp, seed=seed + 1, near_division=near_division, n_var_udf=n_var_udf)
    ns = N_SPECIES[model]
    rng = np.random.default_rng(seed + 2)
    c = conc * (0.5 + rng.random(ns * n_comp))
    flows = []
    if outlet:
        o = n_comp - 1
        q = p_exit * fm["volumes"][o] / dt
        flows = [(o, q, fm["volumes"][o])]
    w = synth.initial_weight(props, x0, float(np.sum(fm["volumes"])))
    return dict(model=model, n=n, n_comp=n_comp, n_species=ns, dt=dt, fm=fm, props=props, pos=pos, conc=c, flows=flows,
                weight=w, n_var_udf=n_var_udf, seed=seed)


def load_case(loop, case, status=None):
    fm = case["fm"]
    loop.set_particles(case["props"], case["pos"], status)
    loop.set_weight(case["weight"])
    if case["n_comp"] > 1:
        loop.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    else:
        loop.domain_update(fm["volumes"], None, fm["out_flows"], None)
    loop.set_leaving_flows(case["flows"])
    loop.set_concentrations(case["conc"])


def conc_at(case, step):
    """deterministic concentration trajectory fed to both implementations (the liquid
    ODE is outside the path: concentrations are a per-step INPUT)"""
    return case["conc"] * (1.0 + 0.05 * np.sin(0.3 * step + np.arange(case["conc"].size)))


def run_steps(loop, case, steps, dt=None, collect=False):
    dt = case["dt"] if dt is None else dt
    srcs = []
    for s in range(steps):
        loop.set_concentrations(conc_at(case, s))
        loop.cycle(dt)
        if collect:
            srcs.append(loop.get_sources().copy())
    return srcs


def assert_state_equal(a, b, n, exact_props=True, rtol=1e-6):
    """a, b: dicts from get_particles"""
    assert np.array_equal(a["position"][:n], b["position"][:n]), "compartment indices differ"
    assert np.array_equal(a["status"][:n], b["status"][:n]), "statuses differ"
    if exact_props:
        assert np.array_equal(a["props"][:, :n].view(np.uint32), b["props"][:, :n].view(np.uint32)), "properties not bit-identical"
        assert np.array_equal(a["age_div"][:n].view(np.uint32), b["age_div"][:n].view(np.uint32))
        assert np.array_equal(a["age_hyd"][:n].view(np.uint32), b["age_hyd"][:n].view(np.uint32))
    else:
        np.testing.assert_allclose(a["props"][:, :n], b["props"][:, :n], rtol=rtol, atol=0)
        np.testing.assert_allclose(a["age_div"][:n], b["age_div"][:n], rtol=rtol, atol=1e-12)
        np.testing.assert_allclose(a["age_hyd"][:n], b["age_hyd"][:n], rtol=rtol, atol=1e-12)


COUNTER_KEYS = ("n_used", "n_inactive", "last_out", "last_dead", "last_waiting_allocation", "total_out", "total_new",
                "n_compactions")


def assert_counters_equal(ca, cb):
    assert ca["events"] == cb["events"], (ca["events"], cb["events"])
    for k in COUNTER_KEYS:
        assert ca[k] == cb[k], (k, ca[k], cb[k])
