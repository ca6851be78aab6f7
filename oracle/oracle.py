"""ctypes binding of the CPU oracle (oracle/libbmc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py.  Same method names as
biocma_mcst_b200.ParticleLoop so parity tests drive both identically.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbmc_oracle.so")
EVENTS = ("NewParticle", "Exit", "Move", "Death", "Overflow", "ChangeWeight")
MODEL_IDS = {"fixed_length": 0, "monod": 1, "simple_acetate": 2, "wide_udf": 3, "udf_model": 4}
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "bmc_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libbmc_oracle.so"], check=True,
                       capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, u64, dbl, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_uint32
        L.orc_create.restype = vp
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_int, u64, u64, u64, u32, ctypes.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_destroy.restype = None
        for n in ("orc_n_used", "orc_capacity", "orc_buffer_capacity", "orc_inactive"):
            getattr(L, n).restype = u64
            getattr(L, n).argtypes = [vp]
        L.orc_n_var.argtypes = [vp]; L.orc_n_c.argtypes = [vp]
        L.orc_set_runtime.argtypes = [vp, u64, dbl, dbl, dbl, dbl]
        L.orc_set_quirk_contrib_return.argtypes = [vp, ctypes.c_int]
        L.orc_set_step.argtypes = [vp, u32]
        L.orc_set_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
        L.orc_get_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
        L.orc_set_weight.argtypes = [vp, dbl]
        L.orc_domain_update.argtypes = [vp, vp, vp, vp, vp, u64]
        L.orc_set_leaving_flows.argtypes = [vp, u64, vp, vp, vp]
        L.orc_set_concentrations.argtypes = [vp, vp]
        L.orc_cycle.argtypes = [vp, dbl]
        L.orc_get_sources.argtypes = [vp, vp]
        L.orc_get_counters.argtypes = [vp, vp]
        L.orc_repartition.argtypes = [vp, vp]
        L.orc_compact.argtypes = [vp]
        L.orc_get_properties.argtypes = [vp, vp, u64, vp, vp, vp, ctypes.POINTER(u64)]
        L.orc_handle_division.argtypes = [vp, u64]
        L.orc_merge_buffer.argtypes = [vp]
        L.orc_set_status.argtypes = [vp, u64, ctypes.c_uint8]
        L.orc_init_particles.argtypes = [vp, u64, ctypes.c_int, vp, ctypes.POINTER(dbl)]
        L.orc_sample.argtypes = [ctypes.c_int, u64, u64, dbl, dbl, dbl, dbl, vp]
        L.orc_ode_step.argtypes = [u64, u64, dbl, vp, vp, vp, vp, vp, u64, vp, vp, vp]
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_philox4x32_10.restype = None
        L.orc_last_error.argtypes = [vp]
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32); k = np.ascontiguousarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return o


def sample(kind, seed, n, p0=0.0, p1=0.0, p2=0.0, p3=0.0):
    kinds = {"normal": 0, "lognormal": 1, "truncated_normal": 2, "truncated_normal_f32": 3, "exponential_f32": 4,
             "drand": 5, "frand": 6, "norminv": 7}
    out = np.empty(n, np.float64)
    rc = lib().orc_sample(kinds[kind], seed, n, p0, p1, p2, p3, _ptr(out))
    assert rc == 0
    return out


def ode_step(C, mass, vol, sink, sources, coo, dt):
    rows, cols, vals = (np.ascontiguousarray(coo[0], np.uint64), np.ascontiguousarray(coo[1], np.uint64),
                        np.ascontiguousarray(coo[2], np.float64))
    ns = C.size // vol.size
    lib().orc_ode_step(ns, vol.size, dt, _ptr(C), _ptr(mass), _ptr(vol), _ptr(sink), _ptr(sources), vals.size,
                       _ptr(rows), _ptr(cols), _ptr(vals))


def ode_step_gl(Cl, ml, vl, sink_l, src_l, coo_l, Cg, mg, vg, sink_g, src_g, coo_g, kla, henry, dt):
    """two-phase ode_step (simulation.model.cpp:131-154); returns the mass-transfer rates of the step"""
    def coo(c):
        return (np.ascontiguousarray(c[0], np.uint64), np.ascontiguousarray(c[1], np.uint64), np.ascontiguousarray(c[2], np.float64))
    rl, cl, valsl = coo(coo_l); rg, cg, valsg = coo(coo_g)
    ns = Cl.size // vl.size
    mtr = np.zeros(Cl.size, np.float64)
    lib().orc_ode_step_gl(ctypes.c_uint64(ns), ctypes.c_uint64(vl.size), ctypes.c_double(dt), _ptr(Cl), _ptr(ml), _ptr(vl), _ptr(sink_l), _ptr(src_l),
                          ctypes.c_uint64(valsl.size), _ptr(rl), _ptr(cl), _ptr(valsl), _ptr(Cg), _ptr(mg), _ptr(vg), _ptr(sink_g), _ptr(src_g),
                          ctypes.c_uint64(valsg.size), _ptr(rg), _ptr(cg), _ptr(valsg), _ptr(np.ascontiguousarray(kla, np.float64)),
                          _ptr(np.ascontiguousarray(henry, np.float64)), _ptr(mtr))
    return mtr


def kla_flowmap_turbulence(ns, eps, vl, vg, db=5e-3):
    """kla of MassTransfer::Type::FlowmapTurbulence (hydro/impl_mtr.cpp:107-140): species 1 from the turbulence correlation"""
    kla = np.zeros(ns * vl.size, np.float64)
    lib().orc_kla_flowmap_turbulence(ctypes.c_uint64(ns), ctypes.c_uint64(vl.size), ctypes.c_double(db), _ptr(np.ascontiguousarray(eps, np.float64)),
                                     _ptr(np.ascontiguousarray(vl, np.float64)), _ptr(np.ascontiguousarray(vg, np.float64)), _ptr(kla))
    return kla


def max_threads():
    return lib().orc_max_threads()


class OracleLoop:
    def __init__(self, model, n_species=1, n_compartments=1, *, seed=2024, rank=0, n_var_udf=32, n_threads=1,
                 allocation_factor=1.5, buffer_ratio=0.6, dead_ratio=0.01, min_removal=0, shrink_ratio=0.0):
        self.L = lib()
        self.model = MODEL_IDS[model] if isinstance(model, str) else int(model)
        self.h = self.L.orc_create(self.model, n_var_udf, n_species, n_compartments, seed, rank, n_threads)
        assert self.h, "orc_create failed"
        self.n_var, self.n_c = self.L.orc_n_var(self.h), self.L.orc_n_c(self.h)
        self.n_species, self.n_compartments = int(n_species), int(n_compartments)
        # same container tuning as the CUDA build of the reference (meson.build:7-12); shrink disabled
        # by default because the device container never shrinks
        self.L.orc_set_runtime(self.h, min_removal, buffer_ratio, allocation_factor, shrink_ratio, dead_ratio)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle error {rc}: {self.L.orc_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_particles(self, props, position=None, status=None, age_hyd=None, age_div=None):
        props = np.ascontiguousarray(props, np.float32)
        n = props.shape[1]
        position = None if position is None else np.ascontiguousarray(position, np.uint64)
        status = None if status is None else np.ascontiguousarray(status, np.uint8)
        age_hyd = None if age_hyd is None else np.ascontiguousarray(age_hyd, np.float32)
        age_div = None if age_div is None else np.ascontiguousarray(age_div, np.float32)
        self._ck(self.L.orc_set_particles(self.h, n, _ptr(props), _ptr(position), _ptr(status), _ptr(age_hyd), _ptr(age_div)))

    def get_particles(self, n=None):
        n = self.L.orc_n_used(self.h) if n is None else int(n)
        props = np.empty((self.n_var, n), np.float32); pos = np.empty(n, np.uint64); st = np.empty(n, np.uint8)
        ah = np.empty(n, np.float32); ad = np.empty(n, np.float32)
        self._ck(self.L.orc_get_particles(self.h, n, _ptr(props), _ptr(pos), _ptr(st), _ptr(ah), _ptr(ad)))
        return dict(props=props, position=pos, status=st, age_hyd=ah, age_div=ad)

    def init_particles(self, n, uniform_position=True, linit=None):
        linit = None if linit is None else np.ascontiguousarray(linit, np.float32)
        m = ctypes.c_double()
        self._ck(self.L.orc_init_particles(self.h, int(n), int(bool(uniform_position)), _ptr(linit), ctypes.byref(m)))
        return m.value

    def set_weight(self, w):
        self.L.orc_set_weight(self.h, float(w))

    def domain_update(self, volumes, neighbors_flat, out_flows, proba_flat):
        vol = np.ascontiguousarray(volumes, np.float64); of = np.ascontiguousarray(out_flows, np.float64)
        if neighbors_flat is None:
            nb = np.zeros((self.n_compartments, 1), np.uint64); pr = np.zeros((self.n_compartments, 1))
        else:
            nb = np.ascontiguousarray(neighbors_flat, np.uint64).reshape(self.n_compartments, -1)
            pr = np.ascontiguousarray(proba_flat, np.float64).reshape(self.n_compartments, -1)
        self._ck(self.L.orc_domain_update(self.h, _ptr(vol), _ptr(nb), _ptr(of), _ptr(pr), nb.shape[1]))

    def set_leaving_flows(self, flows):
        flows = list(flows)
        idx = np.array([f[0] for f in flows], np.uint64); q = np.array([f[1] for f in flows], np.float64)
        v = np.array([f[2] for f in flows], np.float64)
        self._ck(self.L.orc_set_leaving_flows(self.h, len(flows), _ptr(idx), _ptr(q), _ptr(v)))

    def set_concentrations(self, c):
        c = np.ascontiguousarray(c, np.float64)
        self._ck(self.L.orc_set_concentrations(self.h, _ptr(c)))

    def get_sources(self):
        out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.L.orc_get_sources(self.h, _ptr(out)))
        return out

    def cycle(self, d_t):
        self._ck(self.L.orc_cycle(self.h, float(d_t)))

    cycle_process = cycle

    def sync(self):
        pass

    def counters(self):
        c = np.zeros(16, np.uint64)
        self._ck(self.L.orc_get_counters(self.h, _ptr(c)))
        c = [int(x) for x in c]
        return dict(events={EVENTS[i]: c[i] for i in range(6)}, n_used=c[6], n_inactive=c[7], last_out=c[8],
                    last_dead=c[9], last_waiting_allocation=c[10], buffer_index=c[11], capacity=c[12], total_out=c[13],
                    total_new=c[14], n_compactions=c[15], buffer_capacity=int(self.L.orc_buffer_capacity(self.h)))

    def repartition(self):
        out = np.zeros(self.n_compartments, np.uint64)
        self._ck(self.L.orc_repartition(self.h, _ptr(out)))
        return out

    def compact(self):
        self._ck(self.L.orc_compact(self.h))

    def get_properties(self, indices=None, with_age=True):
        idx = None if indices is None else np.ascontiguousarray(indices, np.uint64)
        n_exp = self.n_var if idx is None else idx.size
        n = ctypes.c_uint64()
        self._ck(self.L.orc_get_properties(self.h, _ptr(idx), 0 if idx is None else idx.size, None, None, None, ctypes.byref(n)))
        n_p = n.value
        pv = np.zeros((n_exp + 1, n_p), np.float64); sv = np.zeros((n_exp + 1, self.n_compartments), np.float64)
        ag = np.zeros((2, n_p), np.float64) if with_age else None
        self._ck(self.L.orc_get_properties(self.h, _ptr(idx), 0 if idx is None else idx.size, _ptr(pv), _ptr(sv), _ptr(ag), ctypes.byref(n)))
        return dict(particle_values=pv, spatial_values=sv, ages=ag)

    # hooks used by the container tests (test_container.cpp)
    def handle_division(self, idx):
        return bool(self.L.orc_handle_division(self.h, int(idx)))

    def merge_buffer(self):
        self.L.orc_merge_buffer(self.h)

    def set_status(self, idx, s):
        self.L.orc_set_status(self.h, int(idx), int(s))

    def set_quirk_contrib_return(self, on):
        self.L.orc_set_quirk_contrib_return(self.h, int(bool(on)))
