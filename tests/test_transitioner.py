"""Flow-map schedule (Transitioner) and the reader of the reference's own 0D case directory.

Reference: the main loop rotates flow maps through `d_transionner->need_advance / advance` and `simulation.updateHydro`
(apps/core/src/host_specific.cpp:81-87, 263-266); `compute_n_per_flowmap` (global_initaliser.cpp:103-114); the "n14"
case rotates 14 maps (tools/cases.xml:21-46).  The transitioner class and the case format belong to the un-vendored
rcmtool crate; apps/api/tests/data/0d/ is the one case in the tree."""
import importlib
import os
import struct
import time

import numpy as np
import pytest

import util

# apps/api/tests/data/0d/{cma_case, vofL.raw, flowL.raw}, byte for byte (63 + 12 + 32 bytes; tools/make_golden.py checks
# them against /root/reference when it is mounted)
CMA_CASE_0D = bytes.fromhex("010000000100000001000000" + "00" * 12 + "0200000000000000" + "03" +
                            "0a000000" + b"./vofL.raw".hex() + "00" + "0b000000" + b"./flowL.raw".hex())
VOF_0D = struct.pack("<Id", 1, 0.02)
FLOW_0D = struct.pack("<II", 1, 1) + bytes(24)


def _cma(bmc):
    return importlib.import_module("biocma_mcst_b200.cma")


def _write_case(d, files):
    for name, data in files.items():
        with open(os.path.join(d, name), "wb") as f:
            f.write(data)


def test_embedded_0d_case_is_the_reference_fixture():
    d = "/root/reference/apps/api/tests/data/0d"
    if not os.path.isdir(d):
        pytest.skip("/root/reference not mounted")
    assert open(os.path.join(d, "cma_case"), "rb").read() == CMA_CASE_0D
    assert open(os.path.join(d, "vofL.raw"), "rb").read() == VOF_0D
    assert open(os.path.join(d, "flowL.raw"), "rb").read() == FLOW_0D


def test_read_reference_0d_case(bmc, tmp_path):
    _write_case(tmp_path, {"cma_case": CMA_CASE_0D, "vofL.raw": VOF_0D, "flowL.raw": FLOW_0D})
    fm = _cma(bmc).read_cma_case(str(tmp_path), bmc.cma_build)
    assert fm["files"] == ["./vofL.raw", "./flowL.raw"] and fm["t_per_flowmap"] == 0.0
    assert fm["volumes"].tolist() == [0.02] and fm["out_flows"].tolist() == [0.0]     # one compartment of 0.02 m3, batch
    assert fm["neighbors"].shape == (1, 1) and int(fm["neighbors"][0, 0]) == 0 and fm["cdf"][0, 0] == 0.0


def test_read_case_with_flows(bmc, tmp_path):
    # the same layout with three compartments in a ring: what the reader hands on equals cma_build of the triplets
    vol = np.array([0.01, 0.02, 0.03])
    trip = [(0, 1, 2.0), (1, 2, 2.0), (2, 0, 2.0), (1, 0, 1.0), (0, 1, 1.0)]
    flow = struct.pack("<II", 3, 3) + b"".join(struct.pack("<QQd", *t) for t in trip)
    _write_case(tmp_path, {"cma_case": CMA_CASE_0D, "vofL.raw": struct.pack("<I", 3) + vol.tobytes(), "flowL.raw": flow})
    fm = _cma(bmc).read_cma_case(str(tmp_path), bmc.cma_build)
    want = bmc.cma_build(3, [t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip])
    assert np.array_equal(fm["neighbors"], want["neighbors"]) and np.array_equal(fm["cdf"], want["cdf"])
    assert fm["out_flows"].tolist() == [3.0, 3.0, 2.0] and np.array_equal(fm["volumes"], vol)
    with pytest.raises(ValueError):
        _write_case(tmp_path, {"vofL.raw": struct.pack("<I", 2) + vol.tobytes()})
        _cma(bmc).read_cma_case(str(tmp_path), bmc.cma_build)


def test_transitioner_schedule(bmc, synth):
    cma = _cma(bmc)
    maps = [synth.make_flowmap(16, 0.1, seed=100 + k) for k in range(14)]
    tr = cma.Transitioner(maps, t_per_flow_map=0.5)
    assert tr.size() == 14 and tr.n_per_flowmap(0.1) == 6 and cma.Transitioner(maps[:1], 0.5).n_per_flowmap(0.1) == 1
    assert tr.get_current() is maps[0] and not tr.need_advance(0.0, 0.1) and not tr.need_advance(0.49, 0.1)
    seen, t = [], 0.0
    for _ in range(160):            # 16 s: more than two full rotations
        if tr.need_advance(t, 0.1):
            assert tr.advance(t, 0.1) is maps[tr.index_at(t)]
        seen.append(tr.current)
        t = round(t + 0.1, 10)
    assert seen[:5] == [0] * 5 and seen[5] == 1 and seen[69] == 13 and seen[70] == 0 and set(seen) == set(range(14))
    one = cma.Transitioner(maps[:1], 0.5)
    assert not any(one.need_advance(0.1 * k, 0.1) for k in range(100))
    with pytest.raises(ValueError):
        cma.Transitioner([], 1.0)


@pytest.mark.gpu
def test_fourteen_map_rotation_is_bit_exact(bmc, orc, synth):
    """the "n14" pattern: 14 flow maps rotated every 3 steps, with division, exits and compaction going on; the CUDA
    path follows the oracle bit for bit across every switch, with no host synchronisation added by the switches"""
    cma = _cma(bmc)
    n_comp, dt = 64, 20.0
    case = util.make_case(synth, "monod", 60_000, n_comp, dt=dt, near_division=0.8, p_move=0.3, p_exit=0.3)
    maps = [synth.make_flowmap(n_comp, dt, p_move=0.1 + 0.03 * k, seed=300 + k) for k in range(14)]
    g = bmc.ParticleLoop("monod", 1, n_comp, seed=case["seed"], dead_ratio=0.0005)
    o = orc.OracleLoop("monod", 1, n_comp, seed=case["seed"], n_threads=4, dead_ratio=0.0005)
    util.load_case(g, case); util.load_case(o, case)
    tg, to = cma.Transitioner(maps, 3 * dt), cma.Transitioner(maps, 3 * dt)
    t, switches, host_us = 0.0, 0, []
    for step in range(50):
        for loop, tr in ((g, tg), (o, to)):
            if step == 0 or tr.need_advance(t, dt):
                fm = tr.advance(t, dt)
                t0 = time.perf_counter()
                cma.update_hydro(loop, fm)
                # the outlet follows the volume of its compartment (set_leaving_flow, simulation.model.cpp:101-108)
                oc = n_comp - 1
                loop.set_leaving_flows([(oc, 0.3 * fm["volumes"][oc] / dt, fm["volumes"][oc])])
                if loop is g and step:
                    host_us.append((time.perf_counter() - t0) * 1e6); switches += 1
        g.set_concentrations(util.conc_at(case, step)); g.cycle(dt)
        o.set_concentrations(util.conc_at(case, step)); o.cycle(dt)
        t += dt
        if step % 10 == 9:
            cg, co = g.counters(), o.counters()
            util.assert_counters_equal(cg, co)
            util.assert_state_equal(g.get_particles(co["n_used"]), o.get_particles(co["n_used"]), co["n_used"])
    assert switches == 16 and tg.current == to.current == (49 // 3) % 14
    c = o.counters()
    assert c["total_new"] > 0 and c["total_out"] > 0 and c["n_compactions"] >= 2
    # a switch is a handful of stream-ordered copies from pinned staging: tens of microseconds of host time
    assert np.median(host_us) < 500.0, host_us
    print("flow-map switch, host time per update_hydro [us]: median %.1f max %.1f" % (np.median(host_us), max(host_us)))
