"""C++ host layer (biocma-mcst_b200/host): the reference's container test restated on the host
classes, and the CLI driver's full time loop (feed -> ODE -> cycleProcess, host_specific.cpp:257-313)
against the same loop driven from Python with the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "biocma-mcst_b200", "host")


def test_container_cpp():
    r = subprocess.run([os.path.join(HOST, "test_container")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "test_container OK" in r.stdout, r.stdout + r.stderr


def test_cli_argument_validation():
    exe = os.path.join(HOST, "biocma_b200")
    for args in (["-np", "0", "-d", "1"], ["-np", "10", "-d", "-1"], ["-np", "10", "-d", "1", "-dt", "-0.1"], ["-zz", "1"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 2 and "biocma_b200:" in r.stderr  # cli_parser.cpp:267-289


def _oracle_main_loop(orc, fm, n, d_t, final_time, nex, feed_q, feed_c, x0=0.5):
    """main_loop of host_specific.cpp:215-330 with the oracle as the MC unit"""
    nc = fm["n"]
    o = orc.OracleLoop("fixed_length", 1, nc, seed=2024)
    m_tot = o.init_particles(n, True, None)
    o.set_weight(x0 * float(np.sum(fm["volumes"])) / m_tot)
    o.domain_update(fm["volumes"], fm["neighbors"] if nc > 1 else None, fm["out_flows"], fm["cdf"] if nc > 1 else None)
    vol = np.ascontiguousarray(fm["volumes"], np.float64)
    C = np.ones(nc); mass = C * vol
    sources = np.zeros(nc); sink = np.zeros(nc)
    n_iter = int(final_time / d_t) + 1
    dump_number = min(n_iter, nex) - 1
    dump_interval = n_iter // dump_number + 1 if (nex != 0 and dump_number != 0) else n_iter + 1
    recs = {"C": [], "np": [], "t": []}
    t = 0.0

    def update_feed():
        if feed_q > 0:
            sources[0] += feed_q * feed_c; sink[0] += feed_q
            o.set_leaving_flows([(0, feed_q, vol[0])])
        else:
            o.set_leaving_flows([])

    def dump():
        recs["C"].append(C.copy()); recs["np"].append(o.repartition().copy()); recs["t"].append(t)

    update_feed()
    for it in range(n_iter):
        if nex != 0 and it % dump_interval == 0:
            dump()
        update_feed()
        orc.ode_step(C, mass, vol, sink, sources, fm["coo"], d_t)
        t += d_t
        sources[:] = 0; sink[:] = 0
        o.set_concentrations(C)
        o.cycle(d_t)
        sources[:] = o.get_sources()
    o.compact()
    dump()
    return recs, o.counters()


@pytest.mark.parametrize("n_comp,feed", [(1, 0.0), (1, 2e-4), (64, 2e-4)])
def test_cli_full_loop_matches_oracle_loop(orc, synth, tmp_path, n_comp, feed):
    d_t, final_time, nex, n = 0.05, 2.0, 5, 40_000
    fm = synth.make_flowmap(n_comp, d_t, p_move=0.05)
    case_dir = str(tmp_path / "case")
    synth.write_case(case_dir, fm)
    stem = str(tmp_path / "res")
    args = [os.path.join(HOST, "biocma_b200"), "-np", str(n), "-d", str(final_time), "-dt", str(d_t), "-mn", "fixed_length",
            "-f", case_dir, "-er", stem, "-nex", str(nex), "-feed", str(feed), "-feedc", "5.0", "-nt", "1", "-force", "1"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["balance_ok"] and info["n_compartments"] == n_comp
    recs, co = _oracle_main_loop(orc, fm, n, d_t, final_time, nex, feed, 5.0)
    t = np.fromfile(stem + "_time.raw", np.float64)
    C = np.fromfile(stem + "_concentration_liquid.raw", np.float64).reshape(len(t), n_comp)
    npart = np.fromfile(stem + "_number_particle.raw", np.uint64).reshape(len(t), n_comp)
    assert len(t) == len(recs["t"]) == info["n_records"]
    assert np.allclose(t, recs["t"], rtol=0, atol=1e-12)
    assert np.array_equal(npart, np.array(recs["np"]))                     # occupancy trajectory: exact
    np.testing.assert_allclose(C, np.array(recs["C"]), rtol=1e-9, atol=0)  # concentration trajectory
    assert info["out"] == co["total_out"] and info["new"] == co["total_new"]
    if feed > 0:
        assert info["out"] > 0


def test_cli_serde_resume(synth, tmp_path):
    # -serde (cli_parser.cpp:147-151, serde.cpp:64-219): a run resumed from `<stem>_serde_0.raw` ends where the
    # uninterrupted run of the same number of steps ends.  Batch case (nothing leaves), so that the compactions
    # the exporter forces at dump times — which fall on different steps in the two runs — have nothing to move.
    # The liquid restarts from the archived CONCENTRATIONS (mass = C*V is recomputed, as in the reference).
    d_t, n, n_comp = 20.0, 30_000, 16
    fm = synth.make_flowmap(n_comp, d_t, p_move=0.05)
    case_dir = str(tmp_path / "case")
    synth.write_case(case_dir, fm)
    exe = os.path.join(HOST, "biocma_b200")

    def run(stem, final_time, extra=()):
        args = [exe, "-np", str(n), "-d", str(final_time), "-dt", str(d_t), "-mn", "monod", "-f", case_dir,
                "-er", stem, "-nex", "3"] + list(extra)
        r = subprocess.run(args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        return json.loads(r.stdout.strip().splitlines()[-1])
    a1 = run(str(tmp_path / "a1"), 400.0)                                   # 21 steps, writes a1_serde_0.raw
    assert os.path.getsize(str(tmp_path / "a1_serde_0.raw")) > n * 10
    a2 = run(str(tmp_path / "a2"), 400.0, ["-serde", str(tmp_path / "a1_serde_0.raw")])  # 21 more
    b = run(str(tmp_path / "b"), 820.0)                                    # 42 steps in one go
    assert a1["steps"] == 21 and a2["steps"] == 42 and b["steps"] == 42
    assert b["new"] > 1000 and a1["new"] > 0
    for k in ("n_particles", "new", "out"):
        assert a2[k] == b[k], (k, a2[k], b[k])
    ca = np.fromfile(str(tmp_path / "a2") + "_concentration_liquid.raw", np.float64).reshape(-1, n_comp)[-1]
    cb = np.fromfile(str(tmp_path / "b") + "_concentration_liquid.raw", np.float64).reshape(-1, n_comp)[-1]
    np.testing.assert_allclose(ca, cb, rtol=1e-9, atol=0)
    na = np.fromfile(str(tmp_path / "a2") + "_number_particle.raw", np.uint64).reshape(-1, n_comp)[-1]
    nb = np.fromfile(str(tmp_path / "b") + "_number_particle.raw", np.uint64).reshape(-1, n_comp)[-1]
    assert np.array_equal(na, nb)


def test_cli_reads_the_reference_0d_case_and_initialiser(tmp_path):
    # -f <directory with an rcmtool cma_case> (apps/api/tests/data/0d, written here byte for byte) and -fi <raw f64 file>
    import test_transitioner as tt
    case_dir = tmp_path / "0d"
    case_dir.mkdir()
    tt._write_case(str(case_dir), {"cma_case": tt.CMA_CASE_0D, "vofL.raw": tt.VOF_0D, "flowL.raw": tt.FLOW_0D})
    fi = tmp_path / "c0.raw"
    np.array([3.25]).tofile(str(fi))
    exe = os.path.join(HOST, "biocma_b200")
    stem = str(tmp_path / "res")
    r = subprocess.run([exe, "-np", "20000", "-d", "1.0", "-dt", "0.05", "-mn", "monod", "-f", str(case_dir), "-fi", str(fi), "-er", stem,
                        "-nex", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["balance_ok"] and info["n_compartments"] == 1
    C = np.fromfile(stem + "_concentration_liquid.raw", np.float64)
    assert C[0] == 3.25 and C[-1] < 3.25            # starts from the initialiser, the cells take the substrate up
    bad = tmp_path / "bad.raw"
    np.array([1.0, 2.0]).tofile(str(bad))
    r = subprocess.run([exe, "-np", "20000", "-d", "1.0", "-f", str(case_dir), "-fi", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "-fi" in r.stderr


def test_cli_rotating_flow_maps_and_default_model(tmp_path):
    # rotate:<maps>:<n>:<t_per_flow_map>: the transitioner drives updateHydro inside main_loop (host_specific.cpp:263-266);
    # an unknown model name falls back to the default model with an alert (global_initaliser.cpp:261-271)
    exe = os.path.join(HOST, "biocma_b200")
    stem = str(tmp_path / "rot")
    r = subprocess.run([exe, "-np", "30000", "-d", "8.0", "-dt", "0.1", "-mn", "two_meta_div", "-f", "rotate:14:16:0.5", "-er", stem, "-nex", "4"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "using the default model" in r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["balance_ok"] and info["n_compartments"] == 16 and info["steps"] == 81
    # the exchange rate differs from map to map, so the occupancy keeps being redistributed: every compartment is visited
    npart = np.fromfile(stem + "_number_particle.raw", np.uint64).reshape(-1, 16)
    assert npart[-1].sum() == info["n_particles"] and np.all(npart[-1] > 0)
