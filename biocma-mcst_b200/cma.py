"""Flow-map schedule and case reader (host side of the path: apps/core/src/host_specific.cpp:81-87, 263-266).

`Transitioner` mirrors how the reference's main loop uses `CmaUtils::TransitionnerPtrType` (the class itself lives in
the un-vendored rcmtool crate): a run owns `size()` flow maps, each valid for `t_per_flow_map` seconds and visited in a
loop; before a step the loop asks `need_advance(t, d_t)` and, if so, hands the map returned by `advance(t, d_t)` to
`SimulationUnit::updateHydro` — here `ParticleLoop.domain_update` + `liquid_set_transition`, which are stream-ordered
copies of a few small arrays (no host synchronisation).

`read_cma_case` reads one flow map of an rcmtool case directory as far as the format can be read off the single case in
the reference tree (apps/api/tests/data/0d): `vofL.raw` = u32 n + n f64 volumes, `flowL.raw` = u32 rows, u32 cols, then
{u64 row, u64 col, f64 flow} triplets, `cma_case` = three u32, f64 t_per_flowmap, then the file names as length-prefixed
strings."""
import math
import os
import struct

import numpy as np


class Transitioner:
    def __init__(self, maps, t_per_flow_map):
        if not maps:
            raise ValueError("Transitioner: no flow map")
        if any(m["volumes"].shape != maps[0]["volumes"].shape for m in maps):
            raise ValueError("Transitioner: flow maps of different size")
        self.maps, self.t_per_flow_map, self.current = list(maps), float(t_per_flow_map), 0

    def size(self):
        return len(self.maps)

    def index_at(self, t):
        if len(self.maps) == 1 or not self.t_per_flow_map > 0:
            return 0
        return int(math.floor(t / self.t_per_flow_map)) % len(self.maps)

    def need_advance(self, t, d_t):
        return self.index_at(t) != self.current

    def advance(self, t, d_t):
        self.current = self.index_at(t)
        return self.maps[self.current]

    def get_current(self):
        return self.maps[self.current]

    def n_per_flowmap(self, d_t):
        """compute_n_per_flowmap (global_initaliser.cpp:103-114)"""
        if self.t_per_flow_map == 0 or len(self.maps) == 1:
            return 1
        return int(self.t_per_flow_map / d_t) + 1


def update_hydro(loop, fm, liquid=True):
    """SimulationUnit::updateHydro (simulation.cpp:95-140) on a ParticleLoop / OracleLoop"""
    n = fm["volumes"].size
    if n > 1:
        loop.domain_update(fm["volumes"], fm["neighbors"], fm["out_flows"], fm["cdf"])
    else:
        loop.domain_update(fm["volumes"], None, fm["out_flows"], None)
    if liquid and hasattr(loop, "liquid_set_transition") and "coo" in fm:
        loop.liquid_set_transition(fm["coo"])


def read_cma_case(directory, cma_build):
    """-> dict(volumes, neighbors, out_flows, cdf, coo, t_per_flowmap, files); `cma_build` = biocma_mcst_b200.cma_build"""
    raw = open(os.path.join(directory, "cma_case"), "rb").read()
    if len(raw) < 20:
        raise ValueError("cma_case: file too short")
    header = struct.unpack_from("<III", raw, 0)
    t_per = struct.unpack_from("<d", raw, 12)[0]
    files, o = [], 20
    while o + 4 < len(raw):
        ln = struct.unpack_from("<I", raw, o)[0]
        name = raw[o + 4:o + 4 + ln]
        if 5 <= ln <= 4096 and o + 4 + ln <= len(raw) and name.endswith(b".raw") and all(0x20 <= b < 0x7f for b in name):
            files.append(name.decode())
            o += 4 + ln
        else:
            o += 1
    if len(files) < 2:
        raise ValueError("cma_case: volume / flow file names not found")

    def find(key):
        for f in files:
            if key in f:
                return os.path.join(directory, f)
        raise ValueError(f"cma_case: no {key} file")
    v = open(find("vof"), "rb").read()
    n = struct.unpack_from("<I", v, 0)[0]
    if n == 0 or len(v) != 4 + 8 * n:
        raise ValueError("vofL.raw: size does not match its count")
    vol = np.frombuffer(v, np.float64, n, 4).copy()
    f = open(find("flow"), "rb").read()
    nr, nc = struct.unpack_from("<II", f, 0)
    if nr != n or nc != n or (len(f) - 8) % 24:
        raise ValueError("flowL.raw: shape does not match the volumes")
    trip = np.frombuffer(f, np.dtype([("r", "<u8"), ("c", "<u8"), ("f", "<f8")]), (len(f) - 8) // 24, 8)
    keep = (trip["r"] != trip["c"]) & (trip["f"] > 0)
    b = cma_build(n, trip["r"][keep], trip["c"][keep], trip["f"][keep])
    return dict(volumes=vol, neighbors=b["neighbors"], out_flows=b["out_flows"], cdf=b["cdf"], coo=b["coo"], t_per_flowmap=t_per,
                header=header, files=files)


# ---- gas-liquid mass transfer coefficients (host side; hydro/mass_transfer.cpp, hydro/impl_mtr.cpp) -------------
HENRY_O2 = 3.181e-2          # MassTransferModel's constructor: Henry(1) (hydro/mass_transfer.cpp:113-116)
BUBBLE_DIAMETER = 5e-3       # proxy->db


def c_kinematic_viscosity(temp):
    """water, m2/s (hydro/impl_mtr.cpp:57-83)"""
    if temp < 85:
        p = (0.00000000000282244333 * temp ** 6 - 0.00000000126441088087 * temp ** 5 + 0.00000023336659710795 * temp ** 4
             - 0.0000234079044336466 * temp ** 3 + 0.00144686943485654 * temp ** 2 - 0.0607310297913931 * temp + 1.79194000343777)
    else:
        p = (0.00000000000000178038 * temp ** 6 - 0.00000000000277495333 * temp ** 5 + 0.00000000181964246491 * temp ** 4
             - 0.00000064995487357883 * temp ** 3 + 0.000136367622445752 * temp ** 2 - 0.0166081298727911 * temp + 1.08486933174497)
    return round(p * 0.000001 * 10000000000) / 10000000000


def default_henry(n_species):
    h = np.zeros(n_species)
    if n_species > 1:
        h[1] = HENRY_O2
    return h


def kla_fixed(values, n_comp):
    """Type::FixedKla (FunctorKla, hydro/mass_transfer.cpp:24-38): one value per species, every compartment;
    -> n_species x n_comp, species fastest"""
    v = np.asarray(values, np.float64)
    return np.tile(v, n_comp)


def kla_flowmap_turbulence(n_species, energy_dissipation, liquid_volumes, gas_volumes, db=BUBBLE_DIAMETER, temperature=20.0):
    """Type::FlowmapTurbulence (flowmap_gas_liquid_mass_transfer, hydro/impl_mtr.cpp:107-140): species 1 (oxygen) gets
    kl * a with kl = 0.3 (eps nu)^(1/4) Sc^(-1/2) and the interfacial area a = 6 alpha_g / (db (1 - alpha_g)); re-evaluated
    by the caller on every flow-map change (MassTransferModel::update)"""
    eps = np.asarray(energy_dissipation, np.float64); vl = np.asarray(liquid_volumes, np.float64); vg = np.asarray(gas_volumes, np.float64)
    nu = c_kinematic_viscosity(temperature)
    sc = nu / 1e-9
    alpha = vg / (vl + vg)
    kl = 0.3 * (eps * nu) ** 0.25 * sc ** -0.5
    kla = np.zeros(n_species * vl.size)
    if n_species > 1:
        kla[1::n_species] = kl * (6.0 * alpha / (db * (1 - alpha)))
    return kla
