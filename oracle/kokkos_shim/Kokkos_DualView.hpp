// oracle/kokkos_shim (TEST INFRASTRUCTURE): Kokkos::DualView over the serial stand-in — host and device views are one
// allocation, modify/sync do nothing.  Written from the public API (used by apps/libs/kokkos-eigen/public/kokkos_eigen.hpp).
#pragma once
#include <Kokkos_Core.hpp>
namespace Kokkos {
template <class DataT, class... Props> class DualView {
  // the reference passes `void` for unused slots (KokkosEigen2D<..., exec = void, args = void>): drop them
  template <class... Q> struct Pack {};
  template <class Acc, class... Rest> struct Filter;
  template <class... A> struct Filter<Pack<A...>> { using type = View<DataT, A...>; };
  template <class... A, class P0, class... Rest> struct Filter<Pack<A...>, P0, Rest...> {
    using type = std::conditional_t<std::is_void_v<P0>, typename Filter<Pack<A...>, Rest...>::type, typename Filter<Pack<A..., P0>, Rest...>::type>;
  };
 public:
  using t_host = typename Filter<Pack<>, Props...>::type;
  using t_dev = t_host;
  using t_dev_const = View<std::add_const_t<typename t_host::value_type>**, typename t_host::array_layout>;
  using execution_space = Serial;
  using host_mirror_space = HostSpace;
  using memory_space = HostSpace;
  static_assert(t_host::rank == 2, "kokkos_shim: DualView is provided for rank-2 views");
  DualView() = default;
  DualView(const std::string& label, size_t n0 = 0, size_t n1 = 0) : v_(label, n0, n1) {}
  t_host view_host() const { return v_; }
  t_dev view_device() const { return v_; }
  size_t extent(int k) const { return v_.extent(k); }
  template <class Space> void modify() {}
  template <class Space> void sync() {}
  void modify_host() {}
  void modify_device() {}
  void sync_host() {}
  void sync_device() {}
 private:
  t_host v_;
};
}  // namespace Kokkos
