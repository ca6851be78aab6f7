"""ctypes binding of oracle/_ref/libbmc_ref.so: the REFERENCE'S OWN hot-path sources compiled from
/root/reference over oracle/kokkos_shim (see oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY.  Same method names as oracle.OracleLoop / biocma_mcst_b200.ParticleLoop so one
test body drives all three.  The library can only be (re)built where /root/reference exists; elsewhere the
prebuilt file that travelled with the snapshot is used, and `available()` says whether there is one.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbmc_ref.so")
RELEASE_LIB_PATH = os.path.join(_HERE, "_ref", "libbmc_ref_release.so")  # -O3 -DNDEBUG -ffast-math: timing only
REFERENCE_ROOT = os.environ.get("BMC_REFERENCE_ROOT", "/root/reference")
EVENTS = ("NewParticle", "Exit", "Move", "Death", "Overflow", "ChangeWeight")
MODEL_IDS = {"fixed_length": 0, "monod": 1, "simple_acetate": 2,
             "udf_model": 4}  # the reference's example UDF apps/udf_model/minimal.cpp (checker build only)
RELEASE_MODELS = ("fixed_length", "monod", "simple_acetate")
_lib = None


def can_build():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "apps", "libs", "simulation"))


def build(force=False):
    """compile the reference sources where they lie (oracle/Makefile target `ref`); no-op without them"""
    if not can_build():
        return LIB_PATH if os.path.exists(LIB_PATH) else None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE_ROOT, "ref"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, capture_output=True)
    return LIB_PATH


OWN_TESTS = ("test_container", "test_team_strategy", "test_model_sppecies_name", "test_env_var", "test_rng_2", "test_utils_1", "test_feed",
             "test_load_balancing")
OWN_TESTS_DIR = os.path.join(_HERE, "_ref", "tests")


def build_own_tests():
    """the reference's own unit tests for this path over the shim (oracle/Makefile target `ref_tests`)"""
    if can_build():
        subprocess.run(["make", "-C", _HERE, "-j", "8", "REF=" + REFERENCE_ROOT, "ref_tests"], check=True, capture_output=True)
    return OWN_TESTS_DIR


def available():
    return os.path.exists(LIB_PATH) or can_build()


def _declare(L):
    vp, u64, dbl, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_uint32
    L.ref_create.restype = vp
    L.ref_create.argtypes = [ctypes.c_int, u64, u64, u64, u32, u64]
    L.ref_destroy.argtypes = [vp]; L.ref_destroy.restype = None
    L.ref_last_error.argtypes = [vp]; L.ref_last_error.restype = ctypes.c_char_p
    L.ref_n_var.argtypes = [vp]; L.ref_n_c.argtypes = [vp]
    L.ref_set_runtime.argtypes = [vp, u64, dbl, dbl, dbl, dbl]; L.ref_set_runtime.restype = None
    L.ref_set_step.argtypes = [vp, u32]; L.ref_set_step.restype = None
    L.ref_set_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    L.ref_get_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    L.ref_get_contribs.argtypes = [vp, u64, vp]
    L.ref_set_weight.argtypes = [vp, dbl]; L.ref_set_weight.restype = None
    L.ref_domain_update.argtypes = [vp, vp, vp, vp, vp, u64]
    L.ref_set_leaving_flows.argtypes = [vp, u64, vp, vp, vp]
    L.ref_set_concentrations.argtypes = [vp, vp]
    L.ref_cycle.argtypes = [vp, dbl]
    L.ref_get_sources.argtypes = [vp, vp]
    L.ref_get_counters.argtypes = [vp, vp]
    L.ref_init_particles.argtypes = [vp, u64, ctypes.c_int, vp, ctypes.POINTER(dbl)]
    L.ref_sample.argtypes = [ctypes.c_int, u64, u64, dbl, dbl, dbl, dbl, vp]
    L.ref_get_properties.argtypes = [vp, vp, vp, vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    if hasattr(L, "ref_unit_init"):  # checker build only
        L.ref_unit_init.restype = vp
        L.ref_unit_init.argtypes = [ctypes.c_int, u64, vp, u64, ctypes.c_int, u64, u32, vp, dbl]
        L.ref_unit_destroy.argtypes = [vp]; L.ref_unit_destroy.restype = None
        L.ref_unit_total_mass.argtypes = [vp]; L.ref_unit_total_mass.restype = dbl
        L.ref_unit_weight.argtypes = [vp]; L.ref_unit_weight.restype = dbl
        L.ref_unit_n_particle.argtypes = [vp]; L.ref_unit_n_particle.restype = u64
        L.ref_unit_get.argtypes = [vp, ctypes.c_int, u64, vp, vp, vp]
        L.ref_unit_repartition.argtypes = [vp, vp, u64]
        L.ref_load_tuning_constant.argtypes = [vp]; L.ref_load_tuning_constant.restype = None
        L.ref_uniform_balance.argtypes = [u32, u32, u64]; L.ref_uniform_balance.restype = u64
    L.ref_set_threads.argtypes = [ctypes.c_int]; L.ref_set_threads.restype = None
    L.ref_max_threads.restype = ctypes.c_int
    return L


def lib():
    """the checker build (assertions, IEEE arithmetic, serial unless set_threads is called)"""
    global _lib
    if _lib is None:
        if can_build():
            build()
        _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


_release = None


def release_lib():
    """the timing build: the reference's release flags (meson.build:98-105)"""
    global _release
    if _release is None:
        if can_build():
            build()
        _release = _declare(ctypes.CDLL(RELEASE_LIB_PATH))
    return _release


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


SAMPLE_KINDS = {"normal": 0, "lognormal": 1, "truncated_normal": 2, "truncated_normal_f32": 3, "exponential_f32": 4,
                "drand": 5, "frand": 6, "norminv": 7}


def sample(kind, seed, n, p0=0.0, p1=0.0, p2=0.0, p3=0.0):
    """n draws of a distribution of mc/prng/prng_extension.hpp (same signature as oracle.sample)"""
    out = np.empty(n, np.float64)
    rc = lib().ref_sample(SAMPLE_KINDS[kind], seed, n, p0, p1, p2, p3, _ptr(out))
    assert rc == 0
    return out


def unit_init(model, n, volumes, *, uniform=True, seed=2024, rank=0, linit=None, x0=0.5):
    """The reference's own MC::init<Model> + post_init_weight + getRepartition (mcinit.hpp, mc/src/unit.cpp), for the
    models of its container variant (fixed_length, simple_acetate).  Returns the initialised unit's content."""
    L = lib()
    mid = MODEL_IDS[model]
    n_var = {0: 2, 2: 9}[mid]
    vol = np.ascontiguousarray(volumes, np.float64)
    lin = None if linit is None else np.ascontiguousarray(linit, np.float32)
    h = L.ref_unit_init(mid, int(n), _ptr(vol), vol.size, int(bool(uniform)), int(seed), int(rank), _ptr(lin), float(x0))
    if not h:
        raise RuntimeError("MC::init failed")
    try:
        props = np.empty((n_var, n), np.float32); pos = np.empty(n, np.uint64); w = ctypes.c_float()
        assert L.ref_unit_get(h, mid, int(n), _ptr(props), _ptr(pos), ctypes.byref(w)) == 0
        rep = np.zeros(vol.size, np.uint64)
        assert L.ref_unit_repartition(h, _ptr(rep), vol.size) == 0
        return dict(props=props, position=pos, total_mass=L.ref_unit_total_mass(h), init_weight=L.ref_unit_weight(h),
                    weight_f32=w.value, repartition=rep, n_particle=int(L.ref_unit_n_particle(h)))
    finally:
        L.ref_unit_destroy(h)


def uniform_balance(n_ranks, rank, n):
    """UniformLoadBalancer(n_ranks).balance(rank, n) — apps/core/src/load_balancing"""
    return int(lib().ref_uniform_balance(int(n_ranks), int(rank), int(n)))


def load_tuning_constant():
    """MC::load_tuning_constant (mc/src/unit.cpp:302-343) with the current environment"""
    out = np.zeros(5)
    lib().ref_load_tuning_constant(_ptr(out))
    return dict(minimum_dead_particle_removal=int(out[0]), buffer_ratio=out[1], allocation_factor=out[2], shrink_ratio=out[3],
                dead_particle_ratio_threshold=out[4])


class RefLoop:
    """The reference's cycleProcess on the CPU (serial).  The reference refuses N <= particles_per_team
    (kernels.hpp:130-134,163-167); `particles_per_team` (a power of two, multiple of 32) lowers the
    KernelDispatchOptions so that small cases run."""

    def __init__(self, model, n_species=1, n_compartments=1, *, seed=2024, rank=0, particles_per_team=1024,
                 allocation_factor=1.5, buffer_ratio=0.6, dead_ratio=0.01, min_removal=0, shrink_ratio=0.0,
                 release=False, n_threads=1, **_):
        self.L = release_lib() if release else lib()
        self.L.ref_set_threads(int(n_threads))  # process-wide in that library: leagues / ranges over OpenMP threads
        self.model = MODEL_IDS[model]
        self.h = self.L.ref_create(self.model, n_species, n_compartments, seed, rank, particles_per_team)
        assert self.h, "ref_create failed"
        self.n_var, self.n_c = self.L.ref_n_var(self.h), self.L.ref_n_c(self.h)
        self.n_species, self.n_compartments = int(n_species), int(n_compartments)
        self.L.ref_set_runtime(self.h, min_removal, buffer_ratio, allocation_factor, shrink_ratio, dead_ratio)

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"reference error {rc}: {self.L.ref_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
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
        self._ck(self.L.ref_set_particles(self.h, n, _ptr(props), _ptr(position), _ptr(status), _ptr(age_hyd), _ptr(age_div)))

    def init_particles(self, n, uniform_position=True, linit=None):
        linit = None if linit is None else np.ascontiguousarray(linit, np.float32)
        m = ctypes.c_double()
        self._ck(self.L.ref_init_particles(self.h, int(n), int(bool(uniform_position)), _ptr(linit), ctypes.byref(m)))
        return m.value

    def n_used(self):
        return self.counters()["n_used"]

    def get_particles(self, n=None):
        n = self.n_used() if n is None else int(n)
        props = np.empty((self.n_var, n), np.float32); pos = np.empty(n, np.uint64); st = np.empty(n, np.uint8)
        ah = np.empty(n, np.float32); ad = np.empty(n, np.float32)
        self._ck(self.L.ref_get_particles(self.h, n, _ptr(props), _ptr(pos), _ptr(st), _ptr(ah), _ptr(ad)))
        return dict(props=props, position=pos, status=st, age_hyd=ah, age_div=ad)

    def get_properties(self, indices=None, with_age=True):
        """PostProcessing::get_properties; `indices` is ignored: the model's own get_number() decides (as in the reference)"""
        n, rows = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self.L.ref_get_properties(self.h, None, None, None, ctypes.byref(n), ctypes.byref(rows)))
        pv = np.zeros((rows.value, n.value), np.float64); sv = np.zeros((rows.value, self.n_compartments), np.float64)
        ag = np.zeros((2, n.value), np.float64)
        self._ck(self.L.ref_get_properties(self.h, _ptr(pv), _ptr(sv), _ptr(ag), ctypes.byref(n), ctypes.byref(rows)))
        return dict(particle_values=pv, spatial_values=sv, ages=ag if with_age else None)

    def get_contribs(self, n=None):
        n = self.n_used() if n is None else int(n)
        out = np.empty((self.n_c, n), np.float32)
        self._ck(self.L.ref_get_contribs(self.h, n, _ptr(out)))
        return out

    def set_weight(self, w):
        self.L.ref_set_weight(self.h, float(w))

    def domain_update(self, volumes, neighbors_flat, out_flows, proba_flat):
        vol = np.ascontiguousarray(volumes, np.float64); of = np.ascontiguousarray(out_flows, np.float64)
        if neighbors_flat is None:
            nb = np.zeros((self.n_compartments, 1), np.uint64); pr = np.ones((self.n_compartments, 1))
        else:
            nb = np.ascontiguousarray(neighbors_flat, np.uint64).reshape(self.n_compartments, -1)
            pr = np.ascontiguousarray(proba_flat, np.float64).reshape(self.n_compartments, -1)
        self._ck(self.L.ref_domain_update(self.h, _ptr(vol), _ptr(nb), _ptr(of), _ptr(pr), nb.shape[1]))

    def set_leaving_flows(self, flows):
        flows = list(flows)
        idx = np.array([f[0] for f in flows], np.uint64); q = np.array([f[1] for f in flows], np.float64)
        v = np.array([f[2] for f in flows], np.float64)
        self._ck(self.L.ref_set_leaving_flows(self.h, len(flows), _ptr(idx), _ptr(q), _ptr(v)))

    def set_concentrations(self, c):
        c = np.ascontiguousarray(c, np.float64)
        self._ck(self.L.ref_set_concentrations(self.h, _ptr(c)))

    def get_sources(self):
        out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.L.ref_get_sources(self.h, _ptr(out)))
        return out

    def cycle(self, d_t):
        self._ck(self.L.ref_cycle(self.h, float(d_t)))

    cycle_process = cycle

    def sync(self):
        pass

    def counters(self):
        c = np.zeros(16, np.uint64)
        self._ck(self.L.ref_get_counters(self.h, _ptr(c)))
        c = [int(x) for x in c]
        return dict(events={EVENTS[i]: c[i] for i in range(6)}, n_used=c[6], n_inactive=c[7], last_out=c[8],
                    last_dead=c[9], last_waiting_allocation=c[10], buffer_index=c[11], capacity=c[12], total_out=c[13],
                    total_new=c[14], n_compactions=c[15])


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own liquid / gas scalar solver (oracle/ref_liquid.cpp: implScalar.cpp + hydro/mass_transfer.cpp over
# oracle/eigen_shim, rust_shim, kokkos_shim).  TEST INFRASTRUCTURE.
LIQUID_LIB_PATH = os.path.join(_HERE, "_ref", "libbmc_ref_liquid.so")
_liq = None


def liquid_available():
    return os.path.exists(LIQUID_LIB_PATH) or can_build()


def liquid_lib():
    global _liq
    if _liq is None:
        if not os.path.exists(LIQUID_LIB_PATH):
            build()
        L = ctypes.CDLL(LIQUID_LIB_PATH)
        vp, u64, dbl, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_int
        L.refl_create.restype = vp; L.refl_create.argtypes = [u64, u64, vp]
        L.refl_destroy.restype = None; L.refl_destroy.argtypes = [vp]
        L.refl_enable_gas.argtypes = [vp, vp, vp]
        L.refl_set_kla_henry.argtypes = [vp, vp, vp]
        L.refl_enable_gas_turbulence.argtypes = [vp, vp]
        L.refl_update_mass_transfer.argtypes = [vp, vp, vp, vp, vp]
        L.refl_get_henry.argtypes = [vp, vp]
        L.refl_set_hydro.argtypes = [vp, ci, vp, vp, u64, vp, vp, vp]
        L.refl_set_concentration.argtypes = [vp, ci, vp]
        L.refl_get_concentration.argtypes = [vp, ci, vp]
        L.refl_get_mtr.argtypes = [vp, vp]
        L.refl_step.argtypes = [vp, dbl] + [vp] * 8
        _liq = L
    return _liq


class RefLiquid:
    """ScalarSimulation (liquid, optionally gas + MassTransferModel) of the reference, stepped the way its main loop does:
    update_feed -> ode_step -> clearContribution.  Concentrations are species-fastest flat arrays like everywhere else."""

    def __init__(self, n_species, n_comp, volumes):
        self.L = liquid_lib()
        self.ns, self.nc = int(n_species), int(n_comp)
        v = np.ascontiguousarray(volumes, np.float64)
        self.h = self.L.refl_create(self.ns, self.nc, _ptr(v))
        self.two_phase = False

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refl_destroy(self.h); self.h = None

    def set_hydro(self, volumes, coo, gas=False):
        """updateScalarHydro: volumes, their inverses (rcmtool hands over 1/V) and the transition matrix as COO"""
        v = np.ascontiguousarray(volumes, np.float64)
        inv = np.ascontiguousarray(1.0 / v)
        rows, cols, vals = (np.ascontiguousarray(coo[0], np.uint64), np.ascontiguousarray(coo[1], np.uint64), np.ascontiguousarray(coo[2], np.float64))
        assert self.L.refl_set_hydro(self.h, int(gas), _ptr(v), _ptr(inv), vals.size, _ptr(rows), _ptr(cols), _ptr(vals)) == 0

    def enable_gas(self, gas_volumes, kla_per_species):
        gv = np.ascontiguousarray(gas_volumes, np.float64); k = np.ascontiguousarray(kla_per_species, np.float64)
        assert k.size == self.ns
        assert self.L.refl_enable_gas(self.h, _ptr(gv), _ptr(k)) == 0
        self.two_phase = True

    def enable_gas_turbulence(self, gas_volumes):
        """gas phase + MassTransferModel of Type::FlowmapTurbulence (kla from the turbulence correlation, impl_mtr.cpp)"""
        gv = np.ascontiguousarray(gas_volumes, np.float64)
        assert self.L.refl_enable_gas_turbulence(self.h, _ptr(gv)) == 0
        self.two_phase = True

    def update_mass_transfer(self, liquid_volumes, gas_volumes, energy_dissipation):
        """MassTransferModel::update(state): returns the kla field (species-fastest) the reference derives from the state"""
        vl, vg, eps = (np.ascontiguousarray(x, np.float64) for x in (liquid_volumes, gas_volumes, energy_dissipation))
        out = np.empty(self.ns * self.nc)
        assert self.L.refl_update_mass_transfer(self.h, _ptr(vl), _ptr(vg), _ptr(eps), _ptr(out)) == 0
        return out

    def set_kla_henry(self, kla, henry):
        k = np.ascontiguousarray(kla, np.float64); hh = np.ascontiguousarray(henry, np.float64)
        assert k.size == self.ns * self.nc and hh.size == self.ns
        assert self.L.refl_set_kla_henry(self.h, _ptr(k), _ptr(hh)) == 0

    def default_henry(self):
        out = np.empty(self.ns)
        assert self.L.refl_get_henry(self.h, _ptr(out)) == 0
        return out

    def set_concentration(self, c, gas=False):
        c = np.ascontiguousarray(c, np.float64)
        assert self.L.refl_set_concentration(self.h, int(gas), _ptr(c)) == 0

    def concentration(self, gas=False):
        out = np.empty(self.ns * self.nc)
        assert self.L.refl_get_concentration(self.h, int(gas), _ptr(out)) == 0
        return out

    def mass_transfer(self):
        out = np.empty(self.ns * self.nc)
        assert self.L.refl_get_mtr(self.h, _ptr(out)) == 0
        return out

    def step(self, d_t, mc_sources=None, feeds=(), gas_feeds=()):
        """feeds: dicts {species, input_position, flow, concentration, output_position?, first_of_feed?} (one dict per
        (feed, species) pair like the C ABI's bmc_feed; the sink of a feed is set once, by its first entry)"""
        def pack(fs):
            sp = np.array([f["species"] for f in fs], np.uint64); ip = np.array([f["input_position"] for f in fs], np.uint64)
            val = np.array([f["flow"] * f["concentration"] for f in fs], np.float64)   # set_scalar_feed: fd.flow * concentration
            sk = [f for f in fs if f.get("output_position") is not None and f.get("first_of_feed", 1)]
            sc = np.array([f["output_position"] for f in sk], np.uint64); sf = np.array([f["flow"] for f in sk], np.float64)
            return sp, ip, val, sc, sf

        a, b = pack(list(feeds)), pack(list(gas_feeds))
        n_f = (ctypes.c_uint64 * 2)(a[0].size, b[0].size); n_s = (ctypes.c_uint64 * 2)(a[3].size, b[3].size)
        pairs = [(ctypes.c_void_p * 2)(a[i].ctypes.data if a[i].size else None, b[i].ctypes.data if b[i].size else None) for i in range(5)]
        src = None if mc_sources is None else np.ascontiguousarray(mc_sources, np.float64)
        rc = self.L.refl_step(self.h, float(d_t), _ptr(src), ctypes.addressof(n_f), ctypes.addressof(pairs[0]), ctypes.addressof(pairs[1]),
                              ctypes.addressof(pairs[2]), ctypes.addressof(n_s), ctypes.addressof(pairs[3]), ctypes.addressof(pairs[4]))
        assert rc == 0


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own SimulationUnit stepped by the body of its main loop (oracle/ref_sim.cpp).  TEST INFRASTRUCTURE.
SIM_LIB_PATH = os.path.join(_HERE, "_ref", "libbmc_ref_sim.so")
_sim = None


def sim_available():
    return os.path.exists(SIM_LIB_PATH) or can_build()


def sim_lib():
    global _sim
    if _sim is None:
        if not os.path.exists(SIM_LIB_PATH):
            build()
        L = ctypes.CDLL(SIM_LIB_PATH)
        vp, u64, dbl, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_int
        L.rsim_create.restype = vp
        L.rsim_create.argtypes = [ci, u64, u64, u64, vp, vp, u64, vp, vp, vp, vp, vp]
        L.rsim_destroy.restype = None; L.rsim_destroy.argtypes = [vp]
        L.rsim_last_error.restype = ctypes.c_char_p; L.rsim_last_error.argtypes = [vp]
        L.rsim_set_particles.argtypes = [vp, u64, vp, vp, dbl]
        L.rsim_get_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
        L.rsim_update_hydro.argtypes = [vp, vp, vp, vp, vp, u64, u64, vp, vp, vp]
        L.rsim_step.argtypes = [vp, dbl]
        L.rsim_get_concentrations.argtypes = [vp, vp]
        L.rsim_get_sources.argtypes = [vp, vp]
        L.rsim_get_counters.argtypes = [vp, vp]
        L.rsim_create_two_phase.restype = vp
        L.rsim_create_two_phase.argtypes = [ci, u64, u64, u64, vp, vp, u64, vp, vp, vp, vp, vp, vp, vp, u64, vp, vp, vp, vp, vp, vp]
        L.rsim_set_gas_hydro.argtypes = [vp, vp, u64, vp, vp, vp, vp]
        L.rsim_get_gas.argtypes = [vp, vp, vp]
        L.ref_set_threads.argtypes = [ci]; L.ref_set_threads.restype = None
        _sim = L
    return _sim


class RefSim:
    """Simulation::SimulationUnit of the reference with a liquid phase and constant feeds; `step` is one iteration of the
    reference's main loop: update_feed, ode_step, advance, clearContribution, cycleProcess."""

    def __init__(self, model, n_species, n_comp, volumes, c0, feeds=(), seed=2024, gas=None):
        """gas: None, or dict(volumes, c0, feeds, kla_fixed) for a two-phase unit (kla_fixed None = turbulence correlation)"""
        self.L = sim_lib()
        self.L.ref_set_threads(1)
        self.model = MODEL_IDS[model]
        self.ns, self.nc = int(n_species), int(n_comp)
        v = np.ascontiguousarray(volumes, np.float64); c = np.ascontiguousarray(c0, np.float64)

        def pack(fl):
            return (np.array([f["species"] for f in fl], np.uint64), np.array([f["input_position"] for f in fl], np.uint64),
                    np.array([f["output_position"] for f in fl], np.uint64), np.array([f["flow"] for f in fl], np.float64),
                    np.array([f["concentration"] for f in fl], np.float64))
        fs, fi, fo, ff, fc = pack(feeds)
        if gas is None:
            self.h = self.L.rsim_create(self.model, self.ns, self.nc, seed, _ptr(v), _ptr(c), len(feeds), _ptr(fs), _ptr(fi), _ptr(fo), _ptr(ff), _ptr(fc))
        else:
            gv = np.ascontiguousarray(gas["volumes"], np.float64); g0 = np.ascontiguousarray(gas["c0"], np.float64)
            gs, gi, go, gf, gc = pack(gas.get("feeds", ()))
            kf = None if gas.get("kla_fixed") is None else np.ascontiguousarray(gas["kla_fixed"], np.float64)
            self.h = self.L.rsim_create_two_phase(self.model, self.ns, self.nc, seed, _ptr(v), _ptr(c), len(feeds), _ptr(fs), _ptr(fi), _ptr(fo), _ptr(ff),
                                                  _ptr(fc), _ptr(gv), _ptr(g0), len(gas.get("feeds", ())), _ptr(gs), _ptr(gi), _ptr(go), _ptr(gf), _ptr(gc),
                                                  _ptr(kf))
        assert self.h, "rsim_create failed"
        self.n_var = {0: 2, 1: 6, 2: 9}[self.model]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.rsim_destroy(self.h); self.h = None

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.rsim_last_error(self.h).decode())

    def set_particles(self, props, position, weight):
        props = np.ascontiguousarray(props, np.float32); pos = np.ascontiguousarray(position, np.uint64)
        self._ck(self.L.rsim_set_particles(self.h, props.shape[1], _ptr(props), _ptr(pos), float(weight)))

    def update_hydro(self, fm):
        """SimulationUnit::updateHydro from one flow map of biocma_mcst_b200.synth.make_flowmap"""
        vol = np.ascontiguousarray(fm["volumes"], np.float64); neigh = np.ascontiguousarray(fm["neighbors"], np.uint64)
        proba = np.ascontiguousarray(fm["cdf"], np.float64); out = np.ascontiguousarray(fm["out_flows"], np.float64)
        rows, cols, vals = (np.ascontiguousarray(fm["coo"][0], np.uint64), np.ascontiguousarray(fm["coo"][1], np.uint64),
                            np.ascontiguousarray(fm["coo"][2], np.float64))
        m = neigh.size // vol.size
        self._ck(self.L.rsim_update_hydro(self.h, _ptr(vol), _ptr(neigh), _ptr(proba), _ptr(out), m, vals.size, _ptr(rows), _ptr(cols), _ptr(vals)))

    def set_gas_hydro(self, gas_volumes, coo, energy_dissipation=None):
        """the gas part of the iteration state the next update_hydro hands to SimulationUnit::updateHydro"""
        gv = np.ascontiguousarray(gas_volumes, np.float64)
        rows, cols, vals = (np.ascontiguousarray(coo[0], np.uint64), np.ascontiguousarray(coo[1], np.uint64), np.ascontiguousarray(coo[2], np.float64))
        eps = np.ascontiguousarray(np.ones(self.nc) if energy_dissipation is None else energy_dissipation, np.float64)
        self._ck(self.L.rsim_set_gas_hydro(self.h, _ptr(gv), vals.size, _ptr(rows), _ptr(cols), _ptr(vals), _ptr(eps)))

    def gas(self):
        """(gas concentrations, mass-transfer rates of the last step), species fastest"""
        c = np.empty(self.ns * self.nc); m = np.empty(self.ns * self.nc)
        self._ck(self.L.rsim_get_gas(self.h, _ptr(c), _ptr(m)))
        return c, m

    def step(self, d_t):
        self._ck(self.L.rsim_step(self.h, float(d_t)))

    def concentrations(self):
        out = np.empty(self.ns * self.nc); self._ck(self.L.rsim_get_concentrations(self.h, _ptr(out))); return out

    def sources(self):
        out = np.empty(self.ns * self.nc); self._ck(self.L.rsim_get_sources(self.h, _ptr(out))); return out

    def counters(self):
        c = (ctypes.c_ulonglong * 16)()
        self._ck(self.L.rsim_get_counters(self.h, c))
        return dict(events=dict(zip(EVENTS, c[0:6])), n_used=c[6], n_inactive=c[7], last_out=c[8], last_dead=c[9],
                    last_waiting_allocation=c[10], capacity=c[12], total_out=c[13], total_new=c[14], n_compactions=c[15])

    def get_particles(self, n=None):
        n = self.counters()["n_used"] if n is None else int(n)
        props = np.empty((self.n_var, n), np.float32); pos = np.empty(n, np.uint64); st = np.empty(n, np.uint8)
        ah = np.empty(n, np.float32); ad = np.empty(n, np.float32)
        self._ck(self.L.rsim_get_particles(self.h, n, _ptr(props), _ptr(pos), _ptr(st), _ptr(ah), _ptr(ad)))
        return dict(props=props, position=pos, status=st, age_hyd=ah, age_div=ad)
