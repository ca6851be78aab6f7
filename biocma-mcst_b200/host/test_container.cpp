// The reference's container test (apps/libs/mc/tests/test_container.cpp:10-158) restated on the
// host classes of bmc_host.hpp, running on the GPU through the C ABI.  Counts only, like upstream.
#include <cassert>
#include <cstdio>
#include <memory>
#include <vector>

#include "bmc_host.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  const std::size_t size = 1000;
  {  // basic_test :10-34
    MC::MonteCarloUnit unit(BMC_MODEL_FIXED_LENGTH, 1, 1, 2024);
    std::vector<float> props(2 * size, 1.5e-6f);
    unit.check(bmc_set_particles(unit.handle(), size, props.data(), nullptr, nullptr, nullptr, nullptr));
    CHECK(unit.n_particles() == size);
    CHECK(unit.capacity() == size * 3 / 2);  // capacity() == size * get_allocation_factor() (test_container.cpp:18-19): the LOGICAL extent
    CHECK(unit.get_inactive() == 0);
  }
  {  // div_test / merge_test :56-103 — ndiv mothers at l >= l_max divide in one cycle and are merged
    const std::size_t ndiv = 10;
    MC::MonteCarloUnit unit(BMC_MODEL_FIXED_LENGTH, 1, 1, 2024);
    std::vector<float> props(2 * size);
    for (std::size_t i = 0; i < size; ++i) { props[i] = i < ndiv ? 2.5e-6f : 1.2e-6f; props[size + i] = 2e-6f; }
    unit.check(bmc_set_particles(unit.handle(), size, props.data(), nullptr, nullptr, nullptr, nullptr));
    unit.check(bmc_set_weight(unit.handle(), 1.0));
    const double vol = 0.02, zero = 0.0, c = 1.0;
    unit.check(bmc_domain_update(unit.handle(), &vol, nullptr, &zero, nullptr, 0));
    unit.check(bmc_set_concentrations(unit.handle(), &c));
    unit.check(bmc_cycle(unit.handle(), 1e-3));
    CHECK(unit.n_particles() == ndiv + size);
    CHECK(unit.counters().buffer_index == 0);
    CHECK(unit.get_event<MC::EventType::NewParticle>() == ndiv);
    CHECK(unit.get_event<MC::EventType::Overflow>() == 0);
  }
  {  // clean_test :105-123 and clean_test_and_shrink :125-148
    for (std::size_t to_remove : {std::size_t(10), std::size_t(999)}) {
      MC::MonteCarloUnit unit(BMC_MODEL_FIXED_LENGTH, 1, 1, 2024);
      std::vector<float> props(2 * size, 1.5e-6f);
      std::vector<uint8_t> status(size, 0);
      for (std::size_t i = 0; i < to_remove; ++i) status[(i * 7) % size] = static_cast<uint8_t>(MC::Status::Dead);
      std::size_t dead = 0; for (auto s : status) dead += s != 0;
      unit.check(bmc_set_particles(unit.handle(), size, props.data(), nullptr, status.data(), nullptr, nullptr));
      CHECK(unit.n_particles() == size && unit.get_inactive() == dead);
      unit.force_remove_dead();
      CHECK(unit.n_particles() == size - dead && unit.get_inactive() == 0);
    }
  }
  {  // error behaviour: bad arguments give codes, nothing throws across the ABI
    bmc_ctx* h = nullptr;
    CHECK(bmc_create(&h, nullptr) == BMC_ERR_INVALID);
    CHECK(bmc_cycle(nullptr, 0.1) == BMC_ERR_INVALID);
  }
  std::puts("test_container OK");
  return 0;
}
