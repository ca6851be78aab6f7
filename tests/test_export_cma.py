"""SURVEY.md §8f "next" rows 2 and 3: get_properties export and the native flow-map builder."""
import numpy as np
import pytest

import util


# ----------------------------------------------------------------------------- flow-map builder (host only)
def _triplets_of(fm):
    r, c, v = fm["coo"]
    off = r != c
    return r[off], c[off], v[off]


def test_cma_build_reproduces_the_synthetic_flowmap(bmc, synth):
    """the arrays built natively from the flow matrix equal the ones the synthetic generator derives by hand"""
    for n_comp in (2, 27, 500):
        fm = synth.make_flowmap(n_comp, 0.1, p_move=0.02, seed=7)
        src, dst, flow = _triplets_of(fm)
        b = bmc.cma_build(n_comp, src, dst, flow, n_cols=fm["m"])
        synth.check_flowmap_invariants(b | {"volumes": fm["volumes"]})
        assert b["m"] == fm["m"]
        assert np.array_equal(b["neighbors"], fm["neighbors"])
        np.testing.assert_allclose(b["out_flows"], fm["out_flows"], rtol=1e-14)
        np.testing.assert_allclose(b["cdf"], fm["cdf"], rtol=1e-13, atol=0)
        assert np.all(b["cdf"][np.arange(n_comp), [len(np.unique(row[row != i])) - 1 if np.any(row != i) else 0
                                                    for i, row in enumerate(b["neighbors"])]] == 1.0)
        # transition matrix: columns... every row sums to zero (mass conservation of the Eulerian operator)
        tr, tc, tv = b["coo"]
        rowsum = np.zeros(n_comp); np.add.at(rowsum, tr, tv)
        assert np.max(np.abs(rowsum)) <= 1e-12 * np.max(np.abs(tv))


def test_cma_build_reference_payload_example(bmc):
    """3 compartments x 3 neighbours, probability rows {0, .5, 1} (apps/libs/mpi_w/tests/test_iteration_payload.cpp:19-24)
    come out of flows (0 to the first neighbour, equal flows to the other two) — with zero flows dropped, as a CDF {.5, 1}"""
    src = np.array([0, 0, 1, 1, 2, 2], np.uint64); dst = np.array([1, 2, 0, 2, 0, 1], np.uint64)
    b = bmc.cma_build(3, src, dst, np.full(6, 2.0))
    assert b["m"] == 2 and np.array_equal(b["cdf"], np.tile([0.5, 1.0], (3, 1)))
    assert np.array_equal(b["neighbors"], np.array([[1, 2], [0, 2], [0, 1]], np.uint64))
    assert np.array_equal(b["out_flows"], np.full(3, 4.0))


def test_cma_build_padding_dead_end_and_errors(bmc):
    # compartment 2 has no out-flow: all-zero CDF row, padded with its own index; compartment 0 has one neighbour
    b = bmc.cma_build(3, [0, 1, 1], [1, 0, 2], [1.0, 3.0, 1.0], n_cols=3)
    assert np.array_equal(b["neighbors"], np.array([[1, 0, 0], [0, 2, 1], [2, 2, 2]], np.uint64))
    assert np.array_equal(b["cdf"], np.array([[1, 1, 1], [0.75, 1, 1], [0, 0, 0]], float))
    assert np.array_equal(b["out_flows"], [1.0, 4.0, 0.0])
    with pytest.raises(bmc.BmcError):
        bmc.cma_build(2, [0], [5], [1.0])             # index out of range
    with pytest.raises(bmc.BmcError):
        bmc.cma_build(2, [0], [1], [-1.0])            # negative flow
    with pytest.raises(bmc.BmcError):
        bmc.cma_build(3, [0, 0], [1, 2], [1.0, 1.0], n_cols=1)   # fewer columns than neighbours


# ----------------------------------------------------------------------------- get_properties (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("model,indices", [("monod", None), ("fixed_length", [0]), ("simple_acetate", [0, 2, 4])])
def test_get_properties_matches_oracle(bmc, orc, synth, model, indices):
    case = util.make_case(synth, model, 30_000, 40, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    kw = dict(dead_ratio=0.5)   # inactive particles are still in the container when the export starts: it must compact
    g = bmc.ParticleLoop(model, case["n_species"], 40, seed=case["seed"], **kw)
    o = orc.OracleLoop(model, case["n_species"], 40, seed=case["seed"], n_threads=4, **kw)
    util.load_case(g, case); util.load_case(o, case)
    util.run_steps(g, case, 6); util.run_steps(o, case, 6)
    assert g.counters()["n_inactive"] > 0
    a, b = g.get_properties(indices), o.get_properties(indices)
    assert a["particle_values"].shape == b["particle_values"].shape and a["particle_values"].shape[1] == g.counters()["n_used"]
    assert g.counters()["n_inactive"] == 0                                  # force_remove_dead happened
    if model != "simple_acetate":   # its division draws differ in the last bit (CUDA vs glibc exp/log): compare what does not depend on them
        assert np.array_equal(a["particle_values"], b["particle_values"])
        assert np.array_equal(a["ages"], b["ages"])
        np.testing.assert_allclose(a["spatial_values"], b["spatial_values"], rtol=1e-12)
    else:
        assert np.array_equal(a["particle_values"][0], b["particle_values"][0])        # length
        np.testing.assert_allclose(a["spatial_values"][0], b["spatial_values"][0], rtol=1e-12)
    # per-compartment sums are consistent with the per-particle values and the repartition
    pos = g.get_particles()["position"]
    chk = np.zeros_like(a["spatial_values"])
    for k in range(chk.shape[0]):
        np.add.at(chk[k], pos.astype(np.int64), a["particle_values"][k])
    np.testing.assert_allclose(a["spatial_values"], chk, rtol=1e-12)
