"""Synthetic inputs of the shapes BASELINE.json names.

* flow maps: what rcmtool hands the MC path — volumes, flat neighbours padded
  with the compartment's own index, row-wise cumulative leave probabilities
  ending at 1, total out-flows, COO transition matrix — for a stirred-tank
  lattice (axial x radial x tangential exchange), honouring the invariants of
  apps/libs/cma_utils/tests/test_transport.cpp:42-79.
* particle populations: SoA property columns of the built-in models with the
  reference's initial distributions (monod.hpp:76-78,102-112;
  fixed_length.hpp:109-120).
Pure numpy; used by tests, bench and the C++ driver's Python twin.  Nothing here
touches the oracle.
"""
import numpy as np


def lattice_dims(n):
    """nx*ny*nz == n with the most cubic factorisation (500 -> 5x10x10, 10000 -> 20x20x25)."""
    best = (1, 1, n)
    for a in range(1, int(round(n ** (1 / 3))) + 2):
        if n % a:
            continue
        r = n // a
        for b in range(a, int(r ** 0.5) + 1):
            if r % b == 0:
                c = r // b
                if max(a, b, c) - min(a, b, c) < max(best) - min(best):
                    best = (a, b, c)
    return best


def make_flowmap(n_comp, dt, *, p_move=0.01, v_total=0.02, seed=2024, max_neighbors=6):
    """Stirred-tank lattice flow map.  Flows are symmetric (F_ij = F_ji) so every
    compartment is volume-balanced; they are scaled so that dt * out_flow / V is
    about `p_move` (the auto-dt rule, global_initaliser.cpp:72)."""
    rng = np.random.default_rng(seed)
    if n_comp == 1:
        return dict(n=1, m=1, volumes=np.array([v_total]), neighbors=np.zeros((1, 1), np.uint64),
                    cdf=np.zeros((1, 1)), out_flows=np.zeros(1), coo=(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0)))
    nx, ny, nz = lattice_dims(n_comp)
    idx = np.arange(n_comp).reshape(nx, ny, nz)
    vol = v_total / n_comp * (1.0 + 0.2 * (2 * rng.random(n_comp) - 1))
    F = {}
    def link(a, b):
        if a == b:
            return
        key = (min(a, b), max(a, b))
        if key not in F:
            F[key] = 1.0 + 0.3 * rng.random()
    for x in range(nx):
        for y in range(ny):
            for z in range(nz):
                a = idx[x, y, z]
                if x + 1 < nx:
                    link(a, idx[x + 1, y, z])          # axial, walls at both ends
                if y + 1 < ny:
                    link(a, idx[x, y + 1, z])          # radial
                if nz > 2 or z + 1 < nz:
                    link(a, idx[x, y, (z + 1) % nz])   # tangential, periodic
    nbrs = [[] for _ in range(n_comp)]
    for (a, b), f in sorted(F.items()):
        nbrs[a].append((b, f))
        nbrs[b].append((a, f))
    m = max(max_neighbors, max(len(r) for r in nbrs))
    neighbors = np.empty((n_comp, m), np.uint64)
    cdf = np.ones((n_comp, m), np.float64)
    out = np.zeros(n_comp)
    raw_out = np.array([sum(f for _, f in r) for r in nbrs])
    scale = p_move / dt / np.mean(raw_out / vol)
    rows, cols, vals = [], [], []
    for i, r in enumerate(nbrs):
        neighbors[i, :] = i                         # padding = own index (test_transport.cpp:97-114)
        fl = np.array([f for _, f in r]) * scale
        out[i] = fl.sum()
        cs = np.cumsum(fl) / out[i]
        cs[-1] = 1.0
        for k, (j, _) in enumerate(r):
            neighbors[i, k] = j
            cdf[i, k] = cs[k]
            rows.append(i); cols.append(j); vals.append(fl[k])
        rows.append(i); cols.append(i); vals.append(-out[i])
    return dict(n=n_comp, m=m, volumes=vol, neighbors=neighbors, cdf=cdf, out_flows=out,
                coo=(np.array(rows, np.uint64), np.array(cols, np.uint64), np.array(vals)))


def check_flowmap_invariants(fm):
    """The assertions of the (commented-out) reference test test_transport.cpp:42-79."""
    cdf = fm["cdf"]
    assert np.all(cdf >= 0) and np.all(cdf <= 1)
    assert np.all(np.diff(cdf, axis=1) >= 0)
    last = cdf[:, -1]
    assert np.all((last == 1.0) | (last == 0.0))
    assert np.all(fm["neighbors"] < fm["n"])
    assert np.all(fm["out_flows"] >= 0) and np.all(fm["volumes"] > 0)


def truncated_normal(rng, n, mu, sigma, lo, hi):
    out = np.empty(n)
    filled = 0
    while filled < n:
        x = rng.normal(mu, sigma, size=max(1024, 2 * (n - filled)))
        x = x[(x > lo) & (x < hi)][: n - filled]
        out[filled:filled + x.size] = x
        filled += x.size
    return out


def make_population(model, n, n_comp, *, seed=2024, n_var_udf=32, near_division=0.0):
    """SoA property columns + positions.  `near_division` in [0,1) shifts the length
    distribution towards l_max so that divisions happen within a few steps."""
    rng = np.random.default_rng(seed)
    l_max = np.float32(2e-6)
    lo = 1e-6 + near_division * 0.95e-6
    length = truncated_normal(rng, n, max(1.5e-6, lo + 0.02e-6), 0.375e-6, lo, 2e-6).astype(np.float32)
    pos = rng.integers(0, n_comp, size=n, dtype=np.uint64)
    if model in ("fixed_length", 0, "udf_model", 4):  # the example UDF has fixed_length's property layout
        props = np.stack([length, np.full(n, l_max, np.float32)])
    elif model in ("monod", 1):
        mu_max = np.float32(0.77 / 3600.0)
        K = np.float32((2e-6 / 2.0) / np.log(2.0))
        z = np.zeros(n, np.float32)
        props = np.stack([length, np.full(n, l_max, np.float32), np.full(n, mu_max, np.float32), z,
                          np.full(n, K, np.float32), z])
    elif model in ("simple_acetate", 2):
        a_max = np.float32(2e-6 / 3600.0)
        lmx = truncated_normal(rng, n, 2e-6, 2e-7, 1.4e-6, 2.6e-6).astype(np.float32)
        z = np.zeros(n, np.float32)
        props = np.stack([np.minimum(length, lmx * np.float32(0.98)), lmx, np.full(n, a_max / 2, np.float32),
                          np.full(n, a_max, np.float32), z, z, z, z, z])
    elif model in ("wide_udf", 3):
        cols = [length, np.full(n, l_max, np.float32)]
        for _ in range(n_var_udf - 2):
            cols.append(rng.random(n, dtype=np.float32))
        props = np.stack(cols)
    else:
        raise ValueError(model)
    return np.ascontiguousarray(props, np.float32), pos


def initial_weight(props, x0, v_total, lin_density=None):
    """post_init_weight (mc/src/unit.cpp:232-257): w = X0 * V_tot / m_tot."""
    if lin_density is None:
        lin_density = np.float32(1000.0) * np.float32(np.pi) * np.float32(0.6e-6) * np.float32(0.6e-6) / np.float32(4.0)
    m_tot = float(np.sum(props[0].astype(np.float64) * float(lin_density)))
    return x0 * v_total / m_tot


def write_case(directory, fm):
    """Flat little-endian arrays in the layout the C++ host layer reads (CmaUtils::FlowMap::load)."""
    import os
    os.makedirs(directory, exist_ok=True)
    rows, cols, vals = fm["coo"]
    for name, arr, dt in (("volumes", fm["volumes"], np.float64), ("out_flows", fm["out_flows"], np.float64),
                          ("neighbors", fm["neighbors"], np.uint64), ("proba", fm["cdf"], np.float64),
                          ("transition_rows", rows, np.uint64), ("transition_cols", cols, np.uint64),
                          ("transition_vals", vals, np.float64)):
        np.ascontiguousarray(arr, dt).tofile(os.path.join(directory, name + ".raw"))
