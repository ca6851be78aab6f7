"""Peer-memory all-reduce of the source vector (bmc_p2p_*): the exchange step of the multi-GPU path
(MPI_Reduce of the sources, apps/core/src/sync.cpp:57-78) as a one-shot reduction over peer mappings.

One GPU is enough to check the protocol: the ranks are contexts of this process on the same device, attached by
address (bmc_p2p_attach_local); their kernels run concurrently on the contexts' streams and meet through the same
flags and buffers the IPC-mapped multi-process case uses.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 5])
def test_one_shot_allreduce_sums_in_rank_order(bmc, synth, world):
    n_comp = 200
    case = util.make_case(synth, "simple_acetate", 6_000 * world, n_comp, dt=0.5, p_move=0.2, p_exit=0.2)
    loops = []
    for r in range(world):
        g = bmc.ParticleLoop("simple_acetate", 2, n_comp, seed=case["seed"], rank=r)
        sl = slice(r * 6_000, (r + 1) * 6_000)
        g.set_particles(case["props"][:, sl], case["pos"][sl])
        g.set_weight(case["weight"])
        fm = case["fm"]
        g.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
        g.set_leaving_flows(case["flows"]); g.set_concentrations(case["conc"])
        loops.append(g)
    bases = [g.p2p_region() for g in loops]
    for r, g in enumerate(loops):
        g.p2p_attach_local(world, r, bases)
    for step in range(6):   # several epochs: both buffers and the flag protocol are reused
        for g in loops:
            g.cycle(case["dt"])
        local = [g.get_sources().copy() for g in loops]
        order = list(range(world)) if step % 2 == 0 else list(reversed(range(world)))  # enqueue order must not matter
        for r in order:
            loops[r].allreduce_sources()
        want = np.zeros_like(local[0])
        for s in local:
            want = want + s   # rank order, like the kernel
        for g in loops:
            got = g.get_sources()
            assert np.array_equal(got, want), np.max(np.abs(got - want))
    assert np.any(want != 0)


def test_missing_peer_is_reported_not_hung(bmc, synth):
    # world of 2 where the second rank never calls the all-reduce: the kernel gives up and the next
    # synchronising call reports it (the spin limit is ~10 s of GPU clock)
    case = util.make_case(synth, "monod", 4_000, 8, dt=0.1)
    a = bmc.ParticleLoop("monod", 1, 8, seed=1, rank=0); b = bmc.ParticleLoop("monod", 1, 8, seed=1, rank=1)
    for g in (a, b):
        util.load_case(g, case)
    bases = [a.p2p_region(), b.p2p_region()]
    a.p2p_attach_local(2, 0, bases); b.p2p_attach_local(2, 1, bases)
    a.cycle(case["dt"])
    a.allreduce_sources()
    with pytest.raises(RuntimeError, match="peer"):
        a.counters()
