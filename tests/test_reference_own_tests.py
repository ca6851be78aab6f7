"""The reference's OWN unit tests for this path, compiled as they are (from /root/reference, nothing copied) over
oracle/kokkos_shim and run here — SURVEY.md §8c asks to check the oracle against every fixture the reference's tests
hold; these are those tests themselves, executed against the reference's code on the stand-in the oracle is pinned with:

  apps/libs/mc/tests/test_container.cpp            container sizes, handle_division x10, merge_buffer, remove_inactive, shrink
  apps/libs/mc/tests/test_team_strategy.cpp        every particle visited exactly once for the team / chunk strategies
  apps/libs/mc/tests/test_model_sppecies_name.cpp  species-name extraction of the model concept
  apps/libs/common/tests/test_env_var.cpp          read_env / read_env_or / set_local_env
  apps/libs/models/tests/test_utils_1.cpp          helpers of models/utils.hpp
  apps/libs/simulation/tests/test_feed.cpp         feed laws (constant, step, pulse, exponential) of feed_descriptor.cpp
  apps/core/tests/test_load_balancing.cpp          particle shares of the load balancers (the rule sharding.py restates)
  apps/libs/mc/tests/test_rng_2.cpp                moments (mean, variance, skewness) of Normal, LogNormal, SkewNormal,
                                                   TruncatedNormal (several), Exponential<float>, norminv: 4e7 samples per law,
                                                   5 % tolerance, drawn from the generator the parity tests use (Philox streams
                                                   behind the Kokkos pool interface) through prng_extension.hpp

Built by `make -C oracle ref_tests` (in __graft_entry__.build()); the binaries travel with the snapshot.
"""
import os
import subprocess

import pytest

import ref


def _exe(name):
    path = os.path.join(ref.OWN_TESTS_DIR, name)
    if not os.path.exists(path):
        if not ref.can_build():
            pytest.skip("reference test binaries absent and /root/reference not mounted")
        ref.build_own_tests()
    return path


@pytest.mark.parametrize("name", ["test_container", "test_team_strategy", "test_model_sppecies_name", "test_env_var", "test_utils_1",
                                  "test_feed", "test_load_balancing"])
def test_reference_unit_test_passes_on_the_shim(name, tmp_path):
    r = subprocess.run([_exe(name)], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]


@pytest.mark.skipif(os.environ.get("BMC_SKIP_SLOW") == "1", reason="BMC_SKIP_SLOW=1")
def test_reference_distribution_moments_pass_on_the_shim(tmp_path):
    # ~75 s: 63 moment checks at 4e7 samples each
    r = subprocess.run([_exe("test_rng_2")], cwd=str(tmp_path), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert r.stderr.count("Test  ") >= 60 and "Failed" not in r.stderr


def test_sharding_rule_equals_the_reference_load_balancer(bmc):
    # biocma-mcst_b200/sharding.py against UniformLoadBalancer::balance itself (iload_balancer.cpp:26-49)
    if not ref.available():
        pytest.skip("oracle/_ref/libbmc_ref.so absent and /root/reference not mounted")
    import importlib
    sharding = importlib.import_module("biocma_mcst_b200.sharding")
    for n in (1, 7, 1000, 10**7 + 3, 10**9 + 1, 2 * 10**9 + 5):
        for world in (1, 2, 3, 4, 8):
            got = [sharding.shard_count(n, r, world) for r in range(world)]
            want = [ref.uniform_balance(world, r, n) for r in range(world)]
            assert got == want and sum(got) == n, (n, world, got, want)
