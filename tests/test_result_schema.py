"""The result container carries the reference's dataset names (apps/core/src/dataexporter/*, post_process.cpp:92-170)
and the identities its post-processing checks (post_process.cpp:92-117: sum(number_particle) = new - removed + N0)."""
import numpy as np
import pytest

import util


@pytest.mark.parametrize("model", ["monod", "simple_acetate"])
def test_result_container_has_reference_schema(bmc, orc, synth, model, tmp_path):
    import importlib
    export = importlib.import_module("biocma_mcst_b200.export")
    n, nc = 6000, 12
    case = util.make_case(synth, model, n, nc, near_division=0.5, p_exit=0.2, p_move=0.1, dt=20.0)
    o = orc.OracleLoop(model, case["n_species"], nc, seed=case["seed"])   # any loop object works; the GPU one in test_export_cma.py
    util.load_case(o, case)
    w = export.ResultWriter(model, nc, case["n_species"], number_particles=n, initial_weight=case["weight"],
                            initial_biomass_concentration=0.5, final_time=10 * case["dt"], delta_time=case["dt"])
    t = 0.0
    for s in range(10):
        c = util.conc_at(case, s)
        if s % 5 == 0:
            w.update_fields(t, c, case["fm"]["volumes"]); w.write_particle_dump(o)
        o.set_concentrations(c); o.cycle(case["dt"]); t += case["dt"]
    w.update_fields(t, c, case["fm"]["volumes"]); ex = w.write_particle_dump(o); w.write_final(o, c)
    w.save(str(tmp_path / "result.npz"))
    z = np.load(str(tmp_path / "result.npz"))
    keys = set(z.files)
    names, idx = export.EXPORT_NAMES[model]
    for k in ("records/time", "records/concentration_liquid", "records/volume_liquid", "records/number_particle", "records/tallies",
              "initial_parameters/number_particles", "initial_parameters/delta_time", "misc/n_rank", "misc/species_names",
              "final_result/number_particles", "final_result/events/total_division", "final_result/concentration_liquid"):
        assert k in keys, k
    for k in range(3):
        for name in names + ["mass"]:
            assert f"biological_model/{k}/{name}" in keys and f"biological_model/{k}/spatial/{name}" in keys
        assert f"biological_model/{k}/age" in keys and f"biological_model/{k}/age_hydro" in keys
    assert z["records/number_particle"].shape == (3, nc) and z["records/tallies"].shape == (3, 6)
    assert z["records/concentration_liquid"].shape == (3, nc, case["n_species"])
    co = o.counters()
    # particle balance of the last dump (post_process.cpp:92-117) and consistency of the dump with the container
    assert int(z["records/number_particle"][-1].sum()) == co["total_new"] - co["total_out"] + n == int(z["final_result/number_particles"])
    assert z["biological_model/2/length"].size == co["n_used"] and np.array_equal(z["biological_model/2/mass"], ex["particle_values"][-1])
    assert np.isclose(z["biological_model/2/spatial/mass"].sum(), z["biological_model/2/mass"].sum(), rtol=1e-12)
