"""Generates tests/golden/ref_*.npz: inputs and per-step outputs of the REFERENCE'S OWN hot-path sources
(oracle/_ref/libbmc_ref.so = /root/reference compiled over oracle/kokkos_shim, see oracle/ref_driver.cpp)
on small seeded cases.  Runs only where /root/reference exists; the fixtures are committed so that the
oracle (CPU suite) and the CUDA path (-m gpu suite) are checked against reference outputs everywhere.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref  # noqa: E402
import util  # noqa: E402
from _bmc_loader import load_synth  # noqa: E402

# name -> (model, n, n_comp, steps, particles_per_team, make_case kwargs)
CASES = {
    # stirred-tank lattice, one outlet with a high exit probability: movement, exits, >= 2 compactions, divisions
    "fixed_length_cma": ("fixed_length", 2500, 20, 12, 1024, dict(near_division=0.5, p_exit=0.2, p_move=0.05, dt=20.0)),
    "monod_cma": ("monod", 2500, 20, 12, 1024, dict(near_division=0.5, p_exit=0.2, p_move=0.05, dt=20.0)),
    "simple_acetate_cma": ("simple_acetate", 2500, 20, 12, 1024, dict(near_division=0.5, p_exit=0.2, p_move=0.05, dt=20.0)),
    # no outlet: nothing ever leaves, so the reference's contribution quirk never fires and the
    # source terms are comparable on every step
    "monod_closed": ("monod", 1500, 12, 8, 256, dict(near_division=0.5, outlet=False, p_move=0.1, dt=20.0)),
    "fixed_length_closed": ("fixed_length", 1500, 12, 8, 256, dict(near_division=0.5, outlet=False, p_move=0.1, dt=20.0)),
    "simple_acetate_closed": ("simple_acetate", 1500, 12, 8, 256, dict(near_division=0.5, outlet=False, p_move=0.1, dt=20.0)),
    # 0D: one compartment (Tag0D contribution kernel, no move), chemostat outlet
    "monod_0d": ("monod", 2048 + 77, 1, 10, 1024, dict(near_division=0.5, p_exit=0.05, dt=20.0)),
    # the reference's example user-defined model (apps/udf_model/minimal.cpp behind Models::UdfModel)
    "udf_model_cma": ("udf_model", 2500, 20, 12, 1024, dict(near_division=0.5, p_exit=0.2, p_move=0.05, dt=20.0)),
    "fixed_length_0d_batch": ("fixed_length", 1300, 1, 6, 256, dict(near_division=0.5, outlet=False, dt=20.0)),
    # two outlets (find_flow with n_flows > 1, move_kernel.hpp:105-127)
    "monod_two_outlets": ("monod", 2600, 20, 10, 1024, dict(near_division=0.5, p_exit=0.15, p_move=0.1, dt=20.0, two_outlets=True)),
    # the caller supplies non-zero ages (the CUDA path then keeps them as floats, updated every step)
    "fixed_length_aged": ("fixed_length", 2200, 16, 10, 1024, dict(near_division=0.5, p_exit=0.2, p_move=0.1, dt=20.0, initial_ages=True)),
    # RuntimeParameters: compaction only above 5 % inactive and at least 64 particles (particles_container.hpp:539-557)
    "monod_late_compaction": ("monod", 2400, 20, 14, 1024, dict(near_division=0.5, p_exit=0.3, p_move=0.1, dt=20.0,
                                                               runtime=dict(dead_ratio=0.05, min_removal=64))),
}


# the models' get_number() (partial exports) — None = HasExportPropertiesFull (all properties)
EXPORT_INDICES = {"fixed_length": [0], "monod": [0, 2, 3], "simple_acetate": None, "udf_model": [0]}


def run_case(name, synth):
    model, n, n_comp, steps, ppt, kw = CASES[name]
    kw = dict(kw)
    two_outlets, initial_ages, runtime = kw.pop("two_outlets", False), kw.pop("initial_ages", False), kw.pop("runtime", {})
    case = util.make_case(synth, model, n, n_comp, **kw)
    if two_outlets:  # a second outlet in the middle of the lattice, twice the flow
        o2 = n_comp // 2
        q2 = 2.0 * case["flows"][0][1] * case["fm"]["volumes"][o2] / case["fm"]["volumes"][case["flows"][0][0]]
        case["flows"] = case["flows"] + [(o2, q2, case["fm"]["volumes"][o2])]
    rng = np.random.default_rng(77)
    age_hyd0 = (rng.random(n) * 500.0).astype(np.float32) if initial_ages else np.zeros(n, np.float32)
    age_div0 = (rng.random(n) * 300.0).astype(np.float32) if initial_ages else np.zeros(n, np.float32)
    loop = ref.RefLoop(model, case["n_species"], n_comp, seed=case["seed"], particles_per_team=ppt, **runtime)
    util.load_case(loop, case)
    if initial_ages:
        loop.set_particles(case["props"], case["pos"], None, age_hyd0, age_div0)
        loop.set_weight(case["weight"])
    out = dict(model=model, n=n, n_comp=n_comp, steps=steps, particles_per_team=ppt, dt=case["dt"], seed=case["seed"],
               n_species=case["n_species"], weight=case["weight"], props0=case["props"], pos0=case["pos"].astype(np.uint32),
               conc0=case["conc"], volumes=case["fm"]["volumes"], out_flows=case["fm"]["out_flows"],
               neighbors=np.asarray(case["fm"]["neighbors"], np.uint32), cdf=case["fm"]["cdf"],
               flows=np.array(case["flows"], np.float64).reshape(-1, 3), age_hyd0=age_hyd0, age_div0=age_div0,
               dead_ratio=runtime.get("dead_ratio", 0.01), min_removal=runtime.get("min_removal", 0))
    srcs, counters, inactive_before = [], [], []
    for s in range(steps):
        inactive_before.append(loop.counters()["n_inactive"])
        loop.set_concentrations(util.conc_at(case, s))
        loop.cycle(case["dt"])
        srcs.append(loop.get_sources())
        c = loop.counters()
        counters.append([c["events"][e] for e in ref.EVENTS] + [c[k] for k in util.COUNTER_KEYS])
        if s in (0, steps // 2, steps - 1):
            st = loop.get_particles()
            out[f"props_{s}"] = st["props"]; out[f"pos_{s}"] = st["position"].astype(np.uint32); out[f"status_{s}"] = st["status"]
            out[f"age_hyd_{s}"] = st["age_hyd"]; out[f"age_div_{s}"] = st["age_div"]
    # PostProcessing::get_properties at the end of the run (forces a compaction): what the exporters write
    ex = loop.get_properties()
    out["export_pv"] = ex["particle_values"]; out["export_sv"] = ex["spatial_values"]; out["export_ages"] = ex["ages"]
    out["export_indices"] = np.array(EXPORT_INDICES[model] if EXPORT_INDICES[model] is not None else [-1], np.int64)
    out["snap_steps"] = np.array(sorted({0, steps // 2, steps - 1}))
    out["sources"] = np.array(srcs); out["counters"] = np.array(counters, np.uint64)
    out["inactive_before"] = np.array(inactive_before, np.uint64)
    return out


# kind -> parameters (the sets of apps/libs/mc/tests/test_rng_2.cpp where it has one)
DISTRIBUTIONS = {
    "normal": (1.0, 2.0, 0.0, 0.0), "lognormal": (0.5, 0.3, 0.0, 0.0), "truncated_normal": (1.5e-6, 0.375e-6, 1e-6, 2e-6),
    "truncated_normal_f32": (1.5e-6, 0.375e-6, 1e-6, 2e-6), "exponential_f32": (2.5, 0.0, 0.0, 0.0), "drand": (0.0,) * 4,
    "frand": (0.0,) * 4, "norminv": (0.2, 1.5, 0.0, 0.0),
}
INIT = dict(n=3000, n_comp=37, seed=99)


def run_init(model):
    """MC::init on the reference's model headers: M::init, position = uniform_u(0, n_comp), total mass"""
    ns = 2 if model == "simple_acetate" else 1
    r = ref.RefLoop(model, ns, INIT["n_comp"], seed=INIT["seed"])
    linit = (1e-6 + np.random.default_rng(5).random(INIT["n"]) * 1e-6).astype(np.float32)
    mass = r.init_particles(INIT["n"], True, linit)
    st = r.get_particles(INIT["n"])
    return dict(model=model, linit=linit, mass=mass, props=st["props"], pos=st["position"].astype(np.uint32), **INIT)


def run_distributions():
    return {k: ref.sample(k, 1407, 4096, *p) for k, p in DISTRIBUTIONS.items()}


def main():
    assert ref.can_build(), "needs /root/reference (the fixtures are generated in the build container)"
    ref.build()
    synth = load_synth()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in CASES:
        out = run_case(name, synth)
        path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        c = out["counters"][-1]
        print(f"{name}: n_used {c[6]}, events {c[:6].tolist()}, compactions {c[-1]}, {os.path.getsize(path) / 1024:.0f} KiB")
    for model in ("fixed_length", "monod", "simple_acetate", "udf_model"):
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"refinit_{model}.npz"), **run_init(model))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refdist.npz"), **run_distributions())
    print("init + distributions written")


if __name__ == "__main__":
    main()
