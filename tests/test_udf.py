"""User-defined models (BMC_MODEL_UDF): the reference's `-mn udf_model` + BIOMC_LIB_UDF path
(apps/api/src/udf_handle.cpp:22-39, apps/udf_model/minimal.cpp), here NVRTC-compiled device hooks.

CPU part: the source-level contract compiles for sm_100a without a device, errors carry the
compiler log.  GPU part: the JIT-compiled example model is bit-exact against the oracle's
independent restatement of apps/udf_model/minimal.cpp."""
import os
import textwrap

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UDF_SRC = os.path.join(ROOT, "examples", "minimal_udf.cu")


def test_example_udf_compiles_without_device(bmc):
    ok, log = bmc.udf_check(UDF_SRC)
    assert ok, log


def test_udf_compile_error_is_reported(bmc, tmp_path):
    bad = tmp_path / "bad_udf.cu"
    bad.write_text('#include "bmc_udf.cuh"\nnamespace { constexpr std::size_t nv() { return 2; } }\nthis is not C++;\n')
    ok, log = bmc.udf_check(str(bad))
    assert not ok and "error" in log and "bad_udf.cu" in log
    ok, log = bmc.udf_check(str(tmp_path / "missing.cu"))
    assert not ok and "cannot read" in log


def test_udf_with_random_draws_and_hints_compiles(bmc, tmp_path):
    """a model that uses the generator interface (get_state / drand / normal / free_state) and the
    write-only hint, with 5 properties and 2 contributions"""
    src = tmp_path / "rng_udf.cu"
    src.write_text(textwrap.dedent('''
        #define BMC_UDF_WRITE_ONLY_MASK (1u << 4)
        #include "bmc_udf.cuh"
        namespace {
        using namespace Models;
        enum class pv : uint8_t { length = 0, l_max, pool_a, pool_b, rate, __COUNT__ };
        constexpr std::size_t nvar() { return INDEX_FROM_ENUM(pv::__COUNT__); }
        constexpr std::size_t nc() { return 2; }
        void init(const MC::pool_type& random_pool, std::size_t idx, const UdfModel::SelfParticle& arr, const UdfModel::Config& config) {
          auto gen = random_pool.get_state();
          GET_PROPERTY(pv::length) = config(idx, 0);
          GET_PROPERTY(pv::l_max) = 2e-6f * (0.9f + 0.2f * gen.frand());
          GET_PROPERTY(pv::pool_a) = (float)gen.drand(0., 1.);
          GET_PROPERTY(pv::pool_b) = (float)gen.normal(0.5, 0.1);
          GET_PROPERTY(pv::rate) = 0.f;
          random_pool.free_state(gen);
        }
        MC::Status update(const MC::pool_type& random_pool, float d_t, std::size_t idx, const UdfModel::SelfParticle& arr,
                          const UdfModel::SelfContribs& arr_contribs, const std::size_t position_index, const MC::LocalConcentration& c) {
          const float s = (float)GET_CONCENTRATION(0), a = (float)GET_CONCENTRATION(1);
          const float r = GET_PROPERTY(pv::pool_a) * s / (1e-3f + s) + GET_PROPERTY(pv::pool_b) * a / (1e-4f + a);
          GET_PROPERTY(pv::rate) = r;
          GET_PROPERTY(pv::length) += d_t * 1e-10f * r;
          GET_CONTRIBS(0) = -r; GET_CONTRIBS(1) = 0.1f * r;
          return check_div(GET_PROPERTY(pv::length), GET_PROPERTY(pv::l_max));
        }
        void division(const MC::pool_type& random_pool, std::size_t idx, std::size_t idx2, const MC::DynParticlesModel<float>& arr,
                      const MC::DynParticlesModel<float>& buffer_arr) {
          auto gen = random_pool.get_state();
          GET_PROPERTY(pv::length) /= 2.f;
          COPY_PROPERTY_TO(pv::length, idx2, buffer_arr)
          COPY_PROPERTY_TO(pv::l_max, idx2, buffer_arr)
          GET_PROPERTY_FROM(idx2, buffer_arr, pv::pool_a) = (float)gen.drand();
          GET_PROPERTY_FROM(idx2, buffer_arr, pv::pool_b) = GET_PROPERTY(pv::pool_b);
          GET_PROPERTY_FROM(idx2, buffer_arr, pv::rate) = 0.f;
          random_pool.free_state(gen);
        }
        double mass(std::size_t idx, const UdfModel::SelfParticle& arr) { return GET_PROPERTY(pv::length) * 2.8e-10; }
        }
        EXPORT_MODULE(module, &init, &update, &division, &mass, BMC_UDF_NONE, BMC_UDF_NONE, &nvar, &nc, BMC_UDF_NONE);
    '''))
    ok, log = bmc.udf_check(str(src))
    assert ok, log


def test_oracle_udf_model_differs_from_fixed_length(orc, synth):
    """the example UDF is NOT fixed_length (saturating increment): the restatement must show it"""
    # dt*ldot must be visible next to 1.0 in `d_length / (1.0 + d_t * ldot)`: a (non-physical) huge step
    case_u = util.make_case(synth, "udf_model", 3000, 4, dt=2.0e5, p_move=0.3)
    case_f = dict(case_u, model="fixed_length")
    ou = orc.OracleLoop("udf_model", 1, 4); of = orc.OracleLoop("fixed_length", 1, 4)
    util.load_case(ou, case_u); util.load_case(of, case_f)
    util.run_steps(ou, case_u, 1); util.run_steps(of, case_f, 1)
    pu, pf = ou.get_particles(3000), of.get_particles(3000)
    assert np.array_equal(pu["position"], pf["position"])          # same flow map, same uniforms
    assert not np.array_equal(pu["props"][0], pf["props"][0])      # different growth law
    assert np.all(pu["props"][0] <= pf["props"][0])                # d_length / (1 + dt*ldot) < d_length


# ----------------------------------------------------------------------------- GPU
def _pair(bmc, orc, case, **kw):
    g = bmc.ParticleLoop("udf_model", 1, case["n_comp"], seed=case["seed"], udf_source=UDF_SRC, **kw)
    o = orc.OracleLoop("udf_model", 1, case["n_comp"], seed=case["seed"], n_threads=4, **kw)
    return g, o


def _compare(g, o):
    cg, co = g.counters(), o.counters()
    util.assert_counters_equal(cg, co)
    n = co["n_used"]
    util.assert_state_equal(g.get_particles(n), o.get_particles(n), n)
    assert np.array_equal(g.repartition(), o.repartition())


@pytest.mark.gpu
def test_udf_dims_discovered_at_load(bmc):
    g = bmc.ParticleLoop("udf_model", 1, 8, udf_source=UDF_SRC)
    assert (g.n_var, g.n_c) == (2, 1)   # set_nvar_udf / set_nc_udf (udfmodel_user.cpp:51-56)


@pytest.mark.gpu
def test_udf_selected_by_env_like_the_reference(bmc, monkeypatch):
    monkeypatch.setenv("BIOMC_LIB_UDF", UDF_SRC)
    g = bmc.ParticleLoop("udf_model", 1, 8)
    assert (g.n_var, g.n_c) == (2, 1)
    monkeypatch.delenv("BIOMC_LIB_UDF")
    with pytest.raises(bmc.BmcError):
        bmc.ParticleLoop("udf_model", 1, 8)


@pytest.mark.gpu
def test_udf_bad_source_fails_create(bmc, tmp_path):
    bad = tmp_path / "bad.cu"
    bad.write_text("int x = ;\n")
    with pytest.raises(bmc.BmcError):
        bmc.ParticleLoop("udf_model", 1, 8, udf_source=str(bad))


@pytest.mark.gpu
def test_udf_parity_single_step(bmc, orc, synth):
    case = util.make_case(synth, "udf_model", 50_000, 500, p_move=0.2, p_exit=0.05)
    g, o = _pair(bmc, orc, case)
    util.load_case(g, case); util.load_case(o, case)
    sg = util.run_steps(g, case, 1, collect=True); so = util.run_steps(o, case, 1, collect=True)
    assert np.max(np.abs(sg[0] - so[0])) <= 1e-9 * np.max(np.abs(so[0]))
    _compare(g, o)


@pytest.mark.gpu
def test_udf_parity_division_exit_compaction(bmc, orc, synth):
    case = util.make_case(synth, "udf_model", 100_000, 200, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.3)
    g, o = _pair(bmc, orc, case, dead_ratio=0.0005)
    util.load_case(g, case); util.load_case(o, case)
    for _ in range(3):
        sg = util.run_steps(g, case, 5, collect=True); so = util.run_steps(o, case, 5, collect=True)
        for a, b in zip(sg, so):
            assert np.max(np.abs(a - b)) <= 1e-9 * (np.max(np.abs(b)) + 1e-300)
        _compare(g, o)
    c = g.counters()
    assert c["total_new"] > 0 and c["total_out"] > 0 and c["n_compactions"] >= 2


@pytest.mark.gpu
def test_udf_device_init_matches_oracle(bmc, orc):
    n, nc = 20_000, 16
    linit = (1e-6 + 1e-6 * np.random.default_rng(5).random(n)).astype(np.float32)
    g = bmc.ParticleLoop("udf_model", 1, nc, seed=99, udf_source=UDF_SRC); o = orc.OracleLoop("udf_model", 1, nc, seed=99)
    mg = g.init_particles(n, True, linit); mo = o.init_particles(n, True, linit)
    assert abs(mg - mo) <= 1e-12 * abs(mo)
    util.assert_state_equal(g.get_particles(n), o.get_particles(n), n)


@pytest.mark.gpu
def test_udf_eager_ages_path(bmc, orc, synth):
    """non-zero initial ages select the eager kernel variant of the JIT-compiled model"""
    case = util.make_case(synth, "udf_model", 30_000, 40, dt=20.0, near_division=0.8, p_move=0.3, p_exit=0.2)
    rng = np.random.default_rng(4)
    ah = (10.0 * rng.random(case["n"])).astype(np.float32); ad = (5.0 * rng.random(case["n"])).astype(np.float32)
    g, o = _pair(bmc, orc, case, dead_ratio=0.002)
    util.load_case(g, case); util.load_case(o, case)
    g.set_particles(case["props"], case["pos"], None, ah, ad); o.set_particles(case["props"], case["pos"], None, ah, ad)
    util.run_steps(g, case, 8); util.run_steps(o, case, 8)
    _compare(g, o)
