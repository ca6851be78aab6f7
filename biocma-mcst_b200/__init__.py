"""biocma-mcst_b200 — B200-native Monte-Carlo particle loop of BioCMA-MCST.

Python is only the harness language here (tests, bench, multi-GPU launch through
``torch.distributed``).  The product is ``libbmc_b200.so``: hand-written sm_100a
kernels behind the C ABI of ``include/bmc.h``.  This module binds that ABI with
ctypes, one method per entry point, and fails loudly when the library is
missing — there is no CPU fallback.

The directory name contains a hyphen, so import it through ``load_pkg()`` of
``_bmc_loader.py`` at the repository root (or ``importlib`` directly).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbmc_b200.so")

MODEL_FIXED_LENGTH, MODEL_MONOD, MODEL_SIMPLE_ACETATE, MODEL_WIDE_UDF, MODEL_UDF = 0, 1, 2, 3, 4
MODEL_IDS = {"fixed_length": 0, "monod": 1, "simple_acetate": 2, "wide_udf": 3, "udf_model": 4}
MODEL_DIMS = {0: (2, 1), 1: (6, 1), 2: (9, 2)}  # (n_var, n_c) of the built-in models
STATUS_IDLE, STATUS_DIVISION, STATUS_EXIT, STATUS_DEAD = 0, 1, 2, 3
EVENTS = ("NewParticle", "Exit", "Move", "Death", "Overflow", "ChangeWeight")

# every symbol include/bmc.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "bmc_create", "bmc_destroy", "bmc_last_error", "bmc_model_dims", "bmc_set_particles", "bmc_get_particles",
    "bmc_init_particles", "bmc_set_weight", "bmc_domain_update", "bmc_set_leaving_flows", "bmc_set_concentrations",
    "bmc_get_sources", "bmc_cycle", "bmc_sync", "bmc_get_counters", "bmc_repartition", "bmc_compact", "bmc_reserve",
    "bmc_sources_device", "bmc_concentrations_device", "bmc_stream", "bmc_launch_count", "bmc_kernel_config", "bmc_profile_enable",
    "bmc_profile_read", "bmc_nccl_unique_id", "bmc_comm_init", "bmc_allreduce_sources",
    "bmc_liquid_set_transition", "bmc_liquid_set_feeds", "bmc_liquid_step", "bmc_get_concentrations",
    "bmc_gas_enable", "bmc_gas_update_hydro", "bmc_gas_set_feeds", "bmc_mass_transfer_set", "bmc_get_gas_concentrations",
    "bmc_get_mass_transfer",
    "bmc_p2p_export", "bmc_p2p_attach", "bmc_p2p_region", "bmc_p2p_attach_local", "bmc_p2p_disable",
    "bmc_udf_check", "bmc_get_properties", "bmc_cma_build", "bmc_checkpoint_size", "bmc_checkpoint_save", "bmc_checkpoint_load",
)


class BmcConfig(ctypes.Structure):
    _fields_ = [
        ("device", ctypes.c_int32), ("model", ctypes.c_int32), ("n_var_udf", ctypes.c_int32), ("reserved0", ctypes.c_int32),
        ("n_species", ctypes.c_uint64), ("n_compartments", ctypes.c_uint64), ("capacity", ctypes.c_uint64),
        ("seed", ctypes.c_uint64), ("rank", ctypes.c_uint32), ("reserved1", ctypes.c_uint32),
        ("allocation_factor", ctypes.c_double), ("buffer_ratio", ctypes.c_double),
        ("dead_particle_ratio_threshold", ctypes.c_double), ("shrink_ratio", ctypes.c_double),
        ("minimum_dead_particle_removal", ctypes.c_uint64), ("udf_source_path", ctypes.c_char_p),
    ]


class BmcLeavingFlow(ctypes.Structure):
    _fields_ = [("index", ctypes.c_uint64), ("flow", ctypes.c_double), ("volume", ctypes.c_double)]


class BmcFeed(ctypes.Structure):
    _fields_ = [("species", ctypes.c_uint64), ("input_position", ctypes.c_uint64), ("flow", ctypes.c_double),
                ("concentration", ctypes.c_double), ("output_position", ctypes.c_uint64), ("has_output", ctypes.c_int32),
                ("first_of_feed", ctypes.c_int32)]


class BmcCounters(ctypes.Structure):
    _fields_ = [
        ("events", ctypes.c_uint64 * 6), ("n_used", ctypes.c_uint64), ("n_inactive", ctypes.c_uint64),
        ("last_out", ctypes.c_uint64), ("last_dead", ctypes.c_uint64), ("last_waiting_allocation", ctypes.c_uint64),
        ("buffer_index", ctypes.c_uint64), ("capacity", ctypes.c_uint64), ("total_out", ctypes.c_uint64),
        ("total_new", ctypes.c_uint64), ("n_compactions", ctypes.c_uint64), ("step", ctypes.c_uint64),
        ("buffer_capacity", ctypes.c_uint64), ("physical_capacity", ctypes.c_uint64),
        ("physical_buffer_capacity", ctypes.c_uint64), ("n_reallocations", ctypes.c_uint64),
    ]


class BmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bmc error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen libbmc_b200.so.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("BMC_LIB") or LIB_PATH  # BMC_LIB: tuning builds of the same ABI
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the particle loop).")
    lib = ctypes.CDLL(p)
    vp, u64, dbl = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_double
    P = ctypes.POINTER
    lib.bmc_create.argtypes = [P(vp), P(BmcConfig)]
    lib.bmc_destroy.argtypes = [P(vp)]
    lib.bmc_last_error.argtypes = [vp]
    lib.bmc_last_error.restype = ctypes.c_char_p
    lib.bmc_model_dims.argtypes = [vp, P(ctypes.c_int32), P(ctypes.c_int32)]
    lib.bmc_set_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    lib.bmc_get_particles.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    lib.bmc_init_particles.argtypes = [vp, u64, ctypes.c_int, vp, P(dbl)]
    lib.bmc_set_weight.argtypes = [vp, dbl]
    lib.bmc_domain_update.argtypes = [vp, vp, vp, vp, vp, u64]
    lib.bmc_set_leaving_flows.argtypes = [vp, u64, P(BmcLeavingFlow)]
    lib.bmc_set_concentrations.argtypes = [vp, vp]
    lib.bmc_get_sources.argtypes = [vp, vp]
    lib.bmc_cycle.argtypes = [vp, dbl]
    lib.bmc_sync.argtypes = [vp]
    lib.bmc_get_counters.argtypes = [vp, P(BmcCounters)]
    lib.bmc_repartition.argtypes = [vp, vp]
    lib.bmc_compact.argtypes = [vp]
    lib.bmc_reserve.argtypes = [vp, u64]
    lib.bmc_checkpoint_size.argtypes = [vp, ctypes.POINTER(u64)]
    lib.bmc_checkpoint_save.argtypes = [vp, vp, u64]
    lib.bmc_checkpoint_load.argtypes = [vp, vp, u64]
    lib.bmc_sources_device.argtypes = [vp, P(vp), P(u64)]
    lib.bmc_concentrations_device.argtypes = [vp, P(vp), P(u64)]
    lib.bmc_stream.argtypes = [vp, P(vp)]
    lib.bmc_launch_count.argtypes = [vp, P(u64)]
    lib.bmc_profile_enable.argtypes = [vp, ctypes.c_int]
    lib.bmc_profile_read.argtypes = [vp, P(dbl), P(u64)]
    lib.bmc_nccl_unique_id.argtypes = [vp]
    lib.bmc_comm_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.bmc_allreduce_sources.argtypes = [vp]
    lib.bmc_p2p_export.argtypes = [vp, vp]
    lib.bmc_p2p_attach.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    lib.bmc_p2p_region.argtypes = [vp, P(vp)]
    lib.bmc_p2p_disable.argtypes = [vp]
    lib.bmc_p2p_attach_local.argtypes = [vp, ctypes.c_int, ctypes.c_int, P(vp)]
    lib.bmc_liquid_set_transition.argtypes = [vp, u64, vp, vp, vp]
    lib.bmc_liquid_set_feeds.argtypes = [vp, u64, P(BmcFeed)]
    lib.bmc_liquid_step.argtypes = [vp, dbl]
    lib.bmc_gas_enable.argtypes = [vp, vp, vp]
    lib.bmc_gas_update_hydro.argtypes = [vp, vp, u64, vp, vp, vp]
    lib.bmc_gas_set_feeds.argtypes = [vp, u64, P(BmcFeed)]
    lib.bmc_mass_transfer_set.argtypes = [vp, vp, vp]
    lib.bmc_get_gas_concentrations.argtypes = [vp, vp]
    lib.bmc_get_mass_transfer.argtypes = [vp, vp]
    lib.bmc_get_concentrations.argtypes = [vp, vp]
    lib.bmc_udf_check.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    lib.bmc_get_properties.argtypes = [vp, vp, u64, vp, vp, vp, P(u64)]
    lib.bmc_cma_build.argtypes = [u64, u64, vp, vp, vp, u64, P(u64), vp, vp, vp, vp, vp, vp]
    for name in ABI_SYMBOLS:
        if name != "bmc_last_error":
            getattr(lib, name).restype = ctypes.c_int
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def cma_build(n, src, dst, flow, n_cols=0):
    """Flow matrix triplets -> dict(m, neighbors (n, m), cdf (n, m), out_flows (n), coo=(rows, cols, vals)) (host only)."""
    lib = load_library()
    src = np.ascontiguousarray(src, np.uint64); dst = np.ascontiguousarray(dst, np.uint64); flow = np.ascontiguousarray(flow, np.float64)
    m = ctypes.c_uint64()
    rc = lib.bmc_cma_build(n, flow.size, _ptr(src), _ptr(dst), _ptr(flow), n_cols, ctypes.byref(m), None, None, None, None, None, None)
    if rc != 0:
        raise BmcError(rc, "bmc_cma_build: invalid flow matrix")
    mm = m.value
    nb = np.zeros((n, mm), np.uint64); cdf = np.zeros((n, mm), np.float64); out = np.zeros(n, np.float64)
    nz = int(np.count_nonzero((src != dst) & (flow > 0))) + n
    tr, tc, tv = np.zeros(nz, np.uint64), np.zeros(nz, np.uint64), np.zeros(nz, np.float64)
    rc = lib.bmc_cma_build(n, flow.size, _ptr(src), _ptr(dst), _ptr(flow), mm, ctypes.byref(m), _ptr(nb), _ptr(cdf), _ptr(out),
                           _ptr(tr), _ptr(tc), _ptr(tv))
    if rc != 0:
        raise BmcError(rc, "bmc_cma_build failed")
    return dict(n=n, m=mm, neighbors=nb, cdf=cdf, out_flows=out, coo=(tr, tc, tv))


def udf_check(source_path):
    """Compile-only check of a user model source (NVRTC, sm_100a; no device needed).
    Returns (ok, compiler log or error text)."""
    lib = load_library()
    buf = ctypes.create_string_buffer(1 << 16)
    rc = lib.bmc_udf_check(os.fsencode(source_path), buf, len(buf))
    return rc == 0, buf.value.decode(errors="replace")


class ParticleLoop:
    """One GPU context of the Monte-Carlo particle loop (wraps ``bmc_ctx``).

    Mirrors what ``SimulationUnit`` + ``ParticlesContainer`` + ``ReactorDomain``
    expose to ``cycleProcess`` on the reference (simulation.hpp:183-239).
    """

    def __init__(self, model, n_species=1, n_compartments=1, *, device=0, seed=2024, rank=0, capacity=0,
                 n_var_udf=32, allocation_factor=0.0, buffer_ratio=0.0, dead_ratio=0.0, min_removal=0, udf_source=None,
                 shrink_ratio=0.0):
        self.lib = load_library()
        self.model = MODEL_IDS[model] if isinstance(model, str) else int(model)
        # `-mn udf_model`: source path from the argument, else env BIOMC_LIB_UDF (read by the library)
        self._udf = None if udf_source is None else os.fsencode(udf_source)
        cfg = BmcConfig(device=device, model=self.model, n_var_udf=n_var_udf, n_species=n_species,
                        n_compartments=n_compartments, capacity=capacity, seed=seed, rank=rank,
                        allocation_factor=allocation_factor, buffer_ratio=buffer_ratio,
                        dead_particle_ratio_threshold=dead_ratio, shrink_ratio=shrink_ratio,
                        minimum_dead_particle_removal=min_removal, udf_source_path=self._udf)
        self.h = ctypes.c_void_p()
        rc = self.lib.bmc_create(ctypes.byref(self.h), ctypes.byref(cfg))
        if rc != 0:
            raise BmcError(rc, "bmc_create failed (see stderr)")
        nv, nc = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.lib.bmc_model_dims(self.h, ctypes.byref(nv), ctypes.byref(nc)))
        self.n_var, self.n_c = nv.value, nc.value
        self.n_species, self.n_compartments = int(n_species), int(n_compartments)
        self._keep = []

    def _ck(self, rc):
        if rc != 0:
            raise BmcError(rc, self.lib.bmc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.bmc_destroy(ctypes.byref(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- particle state --------------------------------------------------
    def set_particles(self, props, position=None, status=None, age_hyd=None, age_div=None):
        props = np.ascontiguousarray(props, dtype=np.float32)
        assert props.ndim == 2 and props.shape[0] == self.n_var, "props must be (n_var, n) SoA"
        n = props.shape[1]
        position = None if position is None else np.ascontiguousarray(position, dtype=np.uint64)
        status = None if status is None else np.ascontiguousarray(status, dtype=np.uint8)
        age_hyd = None if age_hyd is None else np.ascontiguousarray(age_hyd, dtype=np.float32)
        age_div = None if age_div is None else np.ascontiguousarray(age_div, dtype=np.float32)
        self._ck(self.lib.bmc_set_particles(self.h, n, _ptr(props), _ptr(position), _ptr(status), _ptr(age_hyd), _ptr(age_div)))

    def get_particles(self, n=None):
        n = self.counters()["n_used"] if n is None else int(n)
        props = np.empty((self.n_var, n), np.float32)
        pos = np.empty(n, np.uint64)
        st = np.empty(n, np.uint8)
        ah = np.empty(n, np.float32)
        ad = np.empty(n, np.float32)
        self._ck(self.lib.bmc_get_particles(self.h, n, _ptr(props), _ptr(pos), _ptr(st), _ptr(ah), _ptr(ad)))
        return dict(props=props, position=pos, status=st, age_hyd=ah, age_div=ad)

    def init_particles(self, n, uniform_position=True, linit=None):
        linit = None if linit is None else np.ascontiguousarray(linit, dtype=np.float32)
        m = ctypes.c_double()
        self._ck(self.lib.bmc_init_particles(self.h, int(n), int(bool(uniform_position)), _ptr(linit), ctypes.byref(m)))
        return m.value

    def set_weight(self, w):
        self._ck(self.lib.bmc_set_weight(self.h, float(w)))

    def reserve(self, capacity):
        self._ck(self.lib.bmc_reserve(self.h, int(capacity)))

    # ---- domain ------------------------------------------------------------
    def domain_update(self, volumes, neighbors_flat, out_flows, proba_flat):
        vol = np.ascontiguousarray(volumes, dtype=np.float64)
        of = np.ascontiguousarray(out_flows, dtype=np.float64)
        if neighbors_flat is None:
            self._ck(self.lib.bmc_domain_update(self.h, _ptr(vol), None, _ptr(of), None, 0))
            return
        nb = np.ascontiguousarray(neighbors_flat, dtype=np.uint64).reshape(self.n_compartments, -1)
        pr = np.ascontiguousarray(proba_flat, dtype=np.float64).reshape(self.n_compartments, -1)
        assert nb.shape == pr.shape
        self._ck(self.lib.bmc_domain_update(self.h, _ptr(vol), _ptr(nb), _ptr(of), _ptr(pr), nb.shape[1]))

    def set_leaving_flows(self, flows):
        """flows: iterable of (index, flow, volume)"""
        flows = list(flows)
        arr = (BmcLeavingFlow * max(1, len(flows)))()
        for i, (idx, q, v) in enumerate(flows):
            arr[i] = BmcLeavingFlow(int(idx), float(q), float(v))
        self._ck(self.lib.bmc_set_leaving_flows(self.h, len(flows), arr))

    # ---- liquid coupling -----------------------------------------------------
    def set_concentrations(self, c):
        if not (isinstance(c, np.ndarray) and c.dtype == np.float64 and c.flags.c_contiguous):
            c = np.ascontiguousarray(c, dtype=np.float64)
        assert c.size == self.n_species * self.n_compartments
        self._ck(self.lib.bmc_set_concentrations(self.h, c.ctypes.data))

    def get_sources(self, out=None):
        if out is None:
            out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.lib.bmc_get_sources(self.h, out.ctypes.data))
        return out

    # ---- liquid phase on the device ---------------------------------------------
    def liquid_set_transition(self, coo):
        rows = np.ascontiguousarray(coo[0], np.uint64); cols = np.ascontiguousarray(coo[1], np.uint64)
        vals = np.ascontiguousarray(coo[2], np.float64)
        self._ck(self.lib.bmc_liquid_set_transition(self.h, vals.size, _ptr(rows), _ptr(cols), _ptr(vals)))

    def liquid_set_feeds(self, feeds):
        """feeds: iterable of dicts(species, input_position, flow, concentration, output_position=None)"""
        feeds = list(feeds)
        arr = (BmcFeed * max(1, len(feeds)))()
        for i, f in enumerate(feeds):
            out = f.get("output_position")
            arr[i] = BmcFeed(int(f.get("species", 0)), int(f["input_position"]), float(f["flow"]), float(f["concentration"]),
                             0 if out is None else int(out), 0 if out is None else 1, int(f.get("first_of_feed", 1)))
        self._ck(self.lib.bmc_liquid_set_feeds(self.h, len(feeds), arr))

    def _feed_array(self, feeds):
        feeds = list(feeds)
        arr = (BmcFeed * max(1, len(feeds)))()
        for i, f in enumerate(feeds):
            out = f.get("output_position")
            arr[i] = BmcFeed(int(f.get("species", 0)), int(f["input_position"]), float(f["flow"]), float(f["concentration"]),
                             0 if out is None else int(out), 0 if out is None else 1, int(f.get("first_of_feed", 1)))
        return len(feeds), arr

    def liquid_step(self, d_t):
        self._ck(self.lib.bmc_liquid_step(self.h, float(d_t)))

    # ---- gas phase: two-phase flow (bmc_gas_*) --------------------------------
    def gas_enable(self, gas_volumes, gas_concentrations=None):
        v = np.ascontiguousarray(gas_volumes, np.float64)
        c = None if gas_concentrations is None else np.ascontiguousarray(gas_concentrations, np.float64)
        self._ck(self.lib.bmc_gas_enable(self.h, _ptr(v), _ptr(c)))

    def gas_update_hydro(self, gas_volumes, coo):
        v = np.ascontiguousarray(gas_volumes, np.float64)
        rows = np.ascontiguousarray(coo[0], np.uint64); cols = np.ascontiguousarray(coo[1], np.uint64)
        vals = np.ascontiguousarray(coo[2], np.float64)
        self._ck(self.lib.bmc_gas_update_hydro(self.h, _ptr(v), vals.size, _ptr(rows), _ptr(cols), _ptr(vals)))

    def gas_set_feeds(self, feeds):
        n, arr = self._feed_array(feeds)
        self._ck(self.lib.bmc_gas_set_feeds(self.h, n, arr))

    def mass_transfer_set(self, kla, henry=None):
        k = np.ascontiguousarray(kla, np.float64)
        h = None if henry is None else np.ascontiguousarray(henry, np.float64)
        self._ck(self.lib.bmc_mass_transfer_set(self.h, _ptr(k), _ptr(h)))

    def get_gas_concentrations(self):
        out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.lib.bmc_get_gas_concentrations(self.h, _ptr(out)))
        return out

    def get_mass_transfer(self):
        out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.lib.bmc_get_mass_transfer(self.h, _ptr(out)))
        return out

    def get_concentrations(self):
        out = np.empty(self.n_species * self.n_compartments, np.float64)
        self._ck(self.lib.bmc_get_concentrations(self.h, _ptr(out)))
        return out

    # ---- hot path ------------------------------------------------------------
    def cycle(self, d_t):
        self._ck(self.lib.bmc_cycle(self.h, float(d_t)))

    cycle_process = cycle  # reference name: SimulationUnit::cycleProcess

    def sync(self):
        self._ck(self.lib.bmc_sync(self.h))

    def counters(self):
        c = BmcCounters()
        self._ck(self.lib.bmc_get_counters(self.h, ctypes.byref(c)))
        d = {k: int(getattr(c, k)) for k, _ in BmcCounters._fields_ if k != "events"}
        d["events"] = {EVENTS[i]: int(c.events[i]) for i in range(6)}
        return d

    def kernel_config(self):
        """launch configuration of the step kernel: which instantiation this context runs"""
        a = (ctypes.c_int32 * 6)()
        self._ck(self.lib.bmc_kernel_config(self.h, a))
        return dict(vec=a[0], block=a[1], block_eager=a[2], grid=a[3], stamped_ages=bool(a[4]), smem_bytes=a[5])

    def repartition(self):
        out = np.empty(self.n_compartments, np.uint64)
        self._ck(self.lib.bmc_repartition(self.h, _ptr(out)))
        return out

    def checkpoint(self):
        """SerDe::save_simulation: the Monte-Carlo unit + concentrations + step counter as bytes"""
        n = ctypes.c_uint64()
        self._ck(self.lib.bmc_checkpoint_size(self.h, ctypes.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._ck(self.lib.bmc_checkpoint_save(self.h, _ptr(buf), ctypes.c_uint64(buf.size)))
        return buf.tobytes()

    def restore(self, blob):
        """SerDe::load_simulation into a context of the same model and dimensions"""
        buf = np.frombuffer(blob, np.uint8)
        self._ck(self.lib.bmc_checkpoint_load(self.h, _ptr(buf), ctypes.c_uint64(buf.size)))

    def compact(self):
        self._ck(self.lib.bmc_compact(self.h))

    def get_properties(self, indices=None, with_age=True):
        """PostProcessing::get_properties: dict(particle_values (n_exp+1, n_p), spatial_values (n_exp+1, n_comp), ages (2, n_p))"""
        idx = None if indices is None else np.ascontiguousarray(indices, np.uint64)
        n_exp = self.n_var if idx is None else idx.size
        n = ctypes.c_uint64()
        self._ck(self.lib.bmc_get_properties(self.h, _ptr(idx), 0 if idx is None else idx.size, None, None, None, ctypes.byref(n)))
        n_p = n.value
        pv = np.zeros((n_exp + 1, n_p), np.float64); sv = np.zeros((n_exp + 1, self.n_compartments), np.float64)
        ag = np.zeros((2, n_p), np.float64) if with_age else None
        self._ck(self.lib.bmc_get_properties(self.h, _ptr(idx), 0 if idx is None else idx.size, _ptr(pv), _ptr(sv), _ptr(ag), ctypes.byref(n)))
        return dict(particle_values=pv, spatial_values=sv, ages=ag)

    # ---- device-side handles -------------------------------------------------
    def sources_device_ptr(self):
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(self.lib.bmc_sources_device(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def concentrations_device_ptr(self):
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(self.lib.bmc_concentrations_device(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def stream_handle(self):
        p = ctypes.c_void_p()
        self._ck(self.lib.bmc_stream(self.h, ctypes.byref(p)))
        return p.value

    def launch_count(self):
        n = ctypes.c_uint64()
        self._ck(self.lib.bmc_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def profile_enable(self, on=True):
        self._ck(self.lib.bmc_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        self._ck(self.lib.bmc_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    # ---- multi GPU -------------------------------------------------------------
    def nccl_unique_id(self):
        buf = np.zeros(128, np.uint8)
        self._ck(self.lib.bmc_nccl_unique_id(_ptr(buf)))
        return buf

    def comm_init(self, n_ranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._ck(self.lib.bmc_comm_init(self.h, int(n_ranks), int(rank), _ptr(uid)))

    def p2p_export(self):
        """64-byte IPC handle of this rank's exchange region (gather them, then p2p_attach)"""
        buf = np.zeros(64, np.uint8)
        self._ck(self.lib.bmc_p2p_export(self.h, _ptr(buf)))
        return buf

    def p2p_attach(self, n_ranks, rank, handles):
        hs = np.ascontiguousarray(handles, np.uint8).reshape(n_ranks * 64)
        self._ck(self.lib.bmc_p2p_attach(self.h, int(n_ranks), int(rank), _ptr(hs)))

    def p2p_disable(self):
        self._ck(self.lib.bmc_p2p_disable(self.h))

    def p2p_region(self):
        base = ctypes.c_void_p()
        self._ck(self.lib.bmc_p2p_region(self.h, ctypes.byref(base)))
        return base.value

    def p2p_attach_local(self, n_ranks, rank, bases):
        arr = (ctypes.c_void_p * n_ranks)(*bases)
        self._ck(self.lib.bmc_p2p_attach_local(self.h, int(n_ranks), int(rank), arr))

    def allreduce_sources(self):
        self._ck(self.lib.bmc_allreduce_sources(self.h))
