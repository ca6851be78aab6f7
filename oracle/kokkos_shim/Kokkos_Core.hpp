// =============================================================================
// oracle/kokkos_shim/Kokkos_Core.hpp — TEST INFRASTRUCTURE, NOT THE PRODUCT.
//
// A minimal, single-threaded stand-in for the part of the Kokkos 5 API that the
// hot-path headers of BioCMA-MCST use.  Its only purpose is to let the REFERENCE'S
// OWN SOURCES (model hooks, cycle / move / leave / contribution functors, the
// particle container, ReactorDomain), compiled where they lie under /root/reference,
// run on the CPU so that oracle/bmc_oracle.cpp and the CUDA path can be checked
// against them (oracle/ref_driver.cpp, oracle/Makefile target `ref`).
//
// What it is NOT: Kokkos.  Teams have one thread; memory spaces are all host; ScatterView is
// duplicated per thread.  By default everything executes serially, in ascending
// index order (the mode every parity test uses: deterministic).  shim::set_threads(n > 1)
// distributes the leagues of a TeamPolicy and the indices of a RangePolicy over n OpenMP
// threads, as the Kokkos OpenMP backend does with team_size 1 — used only to TIME the
// reference's kernels on all host cores (bench.py --impl reference); parallel_scan stays serial.
// Two pieces of Kokkos arithmetic are not reproduced and are re-specified exactly as
// DESIGN.md §2/§4 states for the oracle:
//   * Random_XorShift1024_Pool -> Philox4x32-10 streams selected by shim::rng() (the
//     driver tells the generator which particle / draw block it is serving);
//   * Kokkos::log(float) -> (float)std::log((double)x).
// Build switches: KOKKOS_SHIM_BOUNDS_CHECK (View index checks; ON in the parity library, OFF for the reference's own
// unit tests, like Kokkos' default), KOKKOS_SHIM_XORSHIFT (timing build: an xorshift1024* state per thread instead of
// the Philox streams), KOKKOS_SHIM_OPEN_UNIFORM (frand in (0,1) for the reference's 4e7-sample distribution test).
// Written from the public Kokkos API documentation; contains no Kokkos source.
// =============================================================================
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [ =, *this ]
// a statement block, like Kokkos' own (the reference writes it without a trailing semicolon in places)
#ifdef NDEBUG
#define KOKKOS_ASSERT(...) {}
#else
#define KOKKOS_ASSERT(...) { if (!bool(__VA_ARGS__)) { std::fprintf(stderr, "KOKKOS_ASSERT(%s) failed at %s:%d\n", #__VA_ARGS__, __FILE__, __LINE__); std::abort(); } }
#endif
#define KOKKOS_ENABLE_SERIAL 1

namespace Kokkos {
namespace shim {
int n_threads();            // defined by the driver; 1 = serial (deterministic)
void set_threads(int n);
inline bool in_parallel() {
#ifdef _OPENMP
  return omp_in_parallel();
#else
  return false;
#endif
}
}  // namespace shim

// ---------------------------------------------------------------- spaces, layouts, traits
struct LayoutLeft {
  size_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  LayoutLeft() = default;
  LayoutLeft(size_t a, size_t b = 0, size_t c = 0) { d[0] = a; d[1] = b; d[2] = c; }
};
struct LayoutRight {
  size_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  LayoutRight() = default;
  LayoutRight(size_t a, size_t b = 0, size_t c = 0) { d[0] = a; d[1] = b; d[2] = c; }
};

struct HostSpace { using memory_space = HostSpace; static constexpr const char* name() { return "Host"; } };
using SharedSpace = HostSpace;
using SharedHostPinnedSpace = HostSpace;
struct ScratchMemorySpace {  // bump allocator over the team's scratch buffer
  using memory_space = ScratchMemorySpace;
  mutable char* cur = nullptr;
  char* end = nullptr;
  void* get(size_t bytes) const {
    const size_t a = (bytes + 15) & ~size_t(15);
    if (cur + a > end) throw std::runtime_error("kokkos_shim: team scratch exhausted");
    void* p = cur; cur += a; return p;
  }
};

struct Serial {
  using execution_space = Serial;
  using memory_space = HostSpace;
  using array_layout = LayoutRight;  // host backends: LayoutRight (alias.hpp:52-56: AoS on OpenMP)
  using scratch_memory_space = ScratchMemorySpace;
  using size_type = size_t;
  static constexpr const char* name() { return "Serial(shim)"; }
  void fence() const {}
  void fence(const std::string&) const {}
  static int concurrency() { return 1; }
};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;
using OpenMP = Serial;

template <class A, class B> struct SpaceAccessibility { static constexpr bool accessible = true; static constexpr bool assignable = true; };

enum MemoryTraitsFlags : unsigned { Unmanaged = 1, RandomAccess = 2, Atomic = 4, Restrict = 8, Aligned = 16 };
template <unsigned F = 0> struct MemoryTraits { static constexpr unsigned flags = F; };

struct ALL_t {};
inline constexpr ALL_t ALL{};
struct AUTO_t { constexpr AUTO_t operator()() const { return *this; } };
inline constexpr AUTO_t AUTO{};
struct WithoutInitializing_t {};
inline constexpr WithoutInitializing_t WithoutInitializing{};
struct ViewAllocProp { std::string label; bool init = true; };
inline ViewAllocProp view_alloc(WithoutInitializing_t, const std::string& l) { return {l, false}; }
inline ViewAllocProp view_alloc(const std::string& l) { return {l, true}; }
inline ViewAllocProp view_alloc(const std::string& l, WithoutInitializing_t) { return {l, false}; }

template <class A, class B> using pair = std::pair<A, B>;
template <class A, class B> constexpr std::pair<A, B> make_pair(A a, B b) { return {a, b}; }

template <class T, size_t N> struct Array {
  T m[N > 0 ? N : 1];
  using value_type = T;
  constexpr T& operator[](size_t i) { return m[i]; }
  constexpr const T& operator[](size_t i) const { return m[i]; }
  static constexpr size_t size() { return N; }
  constexpr T* data() { return m; }
  constexpr const T* data() const { return m; }
};

// ---------------------------------------------------------------- data-type analysis
namespace Impl {
template <class T> struct DataType {  // scalar
  using value_type = T;
  static constexpr int n_dyn = 0, n_static = 0;
  static constexpr size_t static_ext(int) { return 0; }
};
template <class T> struct DataType<T*> {
  using value_type = typename DataType<T>::value_type;
  static constexpr int n_dyn = DataType<T>::n_dyn + 1, n_static = DataType<T>::n_static;
  static constexpr size_t static_ext(int k) { return DataType<T>::static_ext(k); }
};
template <class T, size_t N> struct DataType<T[N]> {  // T may itself be a pointer chain: F*[Nd]
  using value_type = typename DataType<T>::value_type;
  static constexpr int n_dyn = DataType<T>::n_dyn, n_static = DataType<T>::n_static + 1;
  static constexpr size_t static_ext(int k) { return k == 0 ? N : DataType<T>::static_ext(k - 1); }
};
// the static extents of `T[A][B]` come outermost first; for F*[Nd] the array layer is outermost and
// holds the LAST dimension.  Dimensions = dynamic ones first, then the static ones in declaration order.

template <class... P> struct PickLayout { using type = void; };
template <class P0, class... P> struct PickLayout<P0, P...> {
  using type = std::conditional_t<std::is_same_v<P0, LayoutLeft> || std::is_same_v<P0, LayoutRight>, P0, typename PickLayout<P...>::type>;
};
template <class... P> struct PickScratch { static constexpr bool value = (std::is_same_v<P, ScratchMemorySpace> || ...); };
}  // namespace Impl

// ---------------------------------------------------------------- View
template <class DataT, class... Props> class View {
  using DTA = Impl::DataType<DataT>;
  using picked_layout = typename Impl::PickLayout<Props...>::type;

 public:
  using data_type = DataT;
  using value_type = typename DTA::value_type;
  using const_value_type = std::add_const_t<value_type>;
  using non_const_value_type = std::remove_const_t<value_type>;
  using array_layout = std::conditional_t<std::is_void_v<picked_layout>, LayoutRight, picked_layout>;
  using execution_space = Serial;
  using memory_space = HostSpace;
  using device_type = Serial;
  using size_type = size_t;
  using HostMirror = View;
  using host_mirror_type = View;
  using pointer_type = value_type*;
  using reference_type = value_type&;
  static constexpr int rank = DTA::n_dyn + DTA::n_static;
  static constexpr int Rank = rank;
  static constexpr int rank_dynamic = DTA::n_dyn;
  static constexpr bool is_left = std::is_same_v<array_layout, LayoutLeft>;

  static constexpr size_t static_extent(unsigned k) { return (int)k < DTA::n_dyn ? 0 : DTA::static_ext((int)k - DTA::n_dyn); }

 private:
  std::shared_ptr<non_const_value_type[]> own_;
  value_type* p_ = nullptr;
  size_t e_[4] = {1, 1, 1, 1};
  std::string label_;

  void set_extents(const size_t* dyn) {
    for (int k = 0; k < 4; ++k) e_[k] = 1;
    for (int k = 0; k < rank; ++k) e_[k] = k < DTA::n_dyn ? dyn[k] : DTA::static_ext(k - DTA::n_dyn);
  }
  void allocate(bool /*init*/) {
    const size_t n = span();
    own_ = std::shared_ptr<non_const_value_type[]>(new non_const_value_type[n > 0 ? n : 1]());
    p_ = own_.get();
  }
  template <class, class...> friend class View;

 public:
  View() { const size_t z[4] = {0, 0, 0, 0}; set_extents(z); }
  // copies made inside a parallel region do not share ownership (Kokkos disables reference counting there too)
  View(const View& o) : own_(shim::in_parallel() ? nullptr : o.own_), p_(o.p_), label_(shim::in_parallel() ? std::string() : o.label_) {
    for (int k = 0; k < 4; ++k) e_[k] = o.e_[k];
  }
  View(View&&) = default;
  View& operator=(const View& o) {
    if (this != &o) { own_ = shim::in_parallel() ? nullptr : o.own_; p_ = o.p_; if (!shim::in_parallel()) label_ = o.label_; for (int k = 0; k < 4; ++k) e_[k] = o.e_[k]; }
    return *this;
  }
  View& operator=(View&&) = default;
  // allocating constructors
  explicit View(const std::string& label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) : label_(label) {
    const size_t d[4] = {n0, n1, n2, 0}; set_extents(d); allocate(true);
  }
  explicit View(const char* label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) : View(std::string(label), n0, n1, n2) {}
  explicit View(const ViewAllocProp& prop, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) : label_(prop.label) {
    const size_t d[4] = {n0, n1, n2, 0}; set_extents(d); allocate(prop.init);
  }
  View(const std::string& label, const array_layout& l) : label_(label) { set_extents(l.d); allocate(true); }
  // wrapping constructors (unmanaged)
  View(value_type* ptr, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) : p_(ptr) { const size_t d[4] = {n0, n1, n2, 0}; set_extents(d); }
  View(value_type* ptr, const array_layout& l) : p_(ptr) { set_extents(l.d); }
  // team scratch
  View(const ScratchMemorySpace& s, size_t n0 = 0, size_t n1 = 0) {
    const size_t d[4] = {n0, n1, 0, 0}; set_extents(d);
    p_ = static_cast<value_type*>(s.get(span() * sizeof(value_type)));
  }
  // converting copy (const-qualification, memory traits, spaces; same rank).  Layouts must agree unless rank <= 1.
  template <class D2, class... P2>
    requires(View<D2, P2...>::rank == rank && std::is_same_v<std::remove_const_t<typename View<D2, P2...>::value_type>, non_const_value_type> &&
             (std::is_const_v<value_type> || !std::is_const_v<typename View<D2, P2...>::value_type>))
  View(const View<D2, P2...>& o) : own_(shim::in_parallel() ? nullptr : o.own_), p_(o.p_), label_(shim::in_parallel() ? std::string() : o.label_) {
    static_assert(rank <= 1 || View<D2, P2...>::is_left == is_left, "kokkos_shim: layout mismatch in View assignment");
    for (int k = 0; k < 4; ++k) e_[k] = o.e_[k];
  }

  constexpr size_t extent(int k) const { return k < 4 ? e_[k] : 1; }
  constexpr int extent_int(int k) const { return (int)extent(k); }
  size_t size() const { return rank == 0 ? 1 : span(); }
  size_t span() const { return e_[0] * e_[1] * e_[2] * e_[3]; }
  value_type* data() const { return p_; }
  const std::string& label() const { return label_; }
  bool is_allocated() const { return p_ != nullptr; }
  int use_count() const { return (int)own_.use_count(); }
  size_t stride(int k) const {
    if (is_left) { size_t s = 1; for (int q = 0; q < k; ++q) s *= e_[q]; return s; }
    size_t s = 1; for (int q = rank - 1; q > k; --q) s *= e_[q]; return s;
  }
  array_layout layout() const { array_layout l; for (int k = 0; k < rank; ++k) l.d[k] = e_[k]; return l; }

  reference_type operator()() const requires(rank == 0) { return p_[0]; }
  template <class I> reference_type operator()(I i) const requires(rank == 1) { return p_[(size_t)i]; }
  template <class I> reference_type operator[](I i) const requires(rank == 1) { return p_[(size_t)i]; }
  template <class I, class J> reference_type operator()(I i, J j) const requires(rank == 2) {
#ifdef KOKKOS_SHIM_BOUNDS_CHECK  // like KOKKOS_ENABLE_DEBUG_BOUNDS_CHECK: on in the checker library, off for the reference's own tests
    if (!((size_t)i < e_[0] && (size_t)j < e_[1])) { std::fprintf(stderr, "kokkos_shim: View '%s' index (%zu,%zu) out of (%zu,%zu)\n", label_.c_str(), (size_t)i, (size_t)j, e_[0], e_[1]); std::abort(); }
#endif
    return is_left ? p_[(size_t)i + e_[0] * (size_t)j] : p_[(size_t)i * e_[1] + (size_t)j];
  }
  template <class I, class J, class K> reference_type operator()(I i, J j, K k) const requires(rank == 3) {
    return is_left ? p_[(size_t)i + e_[0] * ((size_t)j + e_[1] * (size_t)k)] : p_[((size_t)i * e_[1] + (size_t)j) * e_[2] + (size_t)k];
  }

  // shim-internal: re-shape in place (resize / realloc)
  void shim_realloc(const size_t* dyn, bool keep) {
    View old = *this;
    set_extents(dyn); allocate(true);
    if (keep && old.p_) {
      if constexpr (rank == 1) { const size_t n = std::min(old.e_[0], e_[0]); for (size_t i = 0; i < n; ++i) p_[i] = old.p_[i]; }
      else if constexpr (rank == 2) {
        const size_t n0 = std::min(old.e_[0], e_[0]), n1 = std::min(old.e_[1], e_[1]);
        for (size_t i = 0; i < n0; ++i) for (size_t j = 0; j < n1; ++j) (*this)(i, j) = old(i, j);
      }
    }
  }
};

template <class V> struct is_view : std::false_type {};
template <class D, class... P> struct is_view<View<D, P...>> : std::true_type {};
template <class V> inline constexpr bool is_view_v = is_view<V>::value;

template <class D, class... P> void resize(View<D, P...>& v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) {
  const size_t d[4] = {n0, n1, n2, 0};
  bool same = true;
  for (int k = 0; k < View<D, P...>::rank_dynamic; ++k) same = same && v.extent(k) == d[k];
  if (same && v.is_allocated()) return;
  v.shim_realloc(d, true);
}
template <class D, class... P> void realloc(View<D, P...>& v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0) {
  const size_t d[4] = {n0, n1, n2, 0};
  v.shim_realloc(d, false);
}

template <class D, class... P, class S>
  requires(!is_view_v<S> && std::is_convertible_v<S, typename View<D, P...>::non_const_value_type>)
void deep_copy(const View<D, P...>& dst, const S& value) {
  const size_t n = dst.size();
  for (size_t i = 0; i < n; ++i) dst.data()[i] = (typename View<D, P...>::non_const_value_type)value;
}
template <class D1, class... P1, class D2, class... P2> void deep_copy(const View<D1, P1...>& dst, const View<D2, P2...>& src) {
  using A = View<D1, P1...>; using B = View<D2, P2...>;
  static_assert(A::rank == B::rank);
  if constexpr (A::rank <= 1) { const size_t n = std::min(dst.size(), src.size()); for (size_t i = 0; i < n; ++i) dst.data()[i] = src.data()[i]; }
  else if constexpr (A::rank == 2) {
    for (size_t i = 0; i < dst.extent(0); ++i) for (size_t j = 0; j < dst.extent(1); ++j) dst(i, j) = src(i, j);
  }
}
template <class Space, class D1, class... P1, class D2, class... P2> void deep_copy(const Space&, const View<D1, P1...>& dst, const View<D2, P2...>& src) {
  deep_copy(dst, src);
}
template <class Space, class D, class... P> View<D, P...> create_mirror_view_and_copy(const Space&, const View<D, P...>& v, const std::string& = "") { return v; }
template <class D, class... P> View<D, P...> create_mirror_view(const View<D, P...>& v) { return v; }
template <class Space, class D, class... P> View<D, P...> create_mirror_view(const Space&, const View<D, P...>& v) { return v; }

// subview: the two forms the reference uses — (view2d, ALL, j) and (view1d, pair)
template <class T> struct StridedColumn {  // column j of a rank-2 view, any layout
  T* p = nullptr; size_t n = 0, s = 1;
  using value_type = T;
  static constexpr int rank = 1;
  T& operator()(size_t i) const { return p[i * s]; }
  T& operator[](size_t i) const { return p[i * s]; }
  size_t extent(int k) const { return k == 0 ? n : 1; }
  size_t size() const { return n; }
  T* data() const { return p; }
};
template <class D, class... P, class J> auto subview(const View<D, P...>& v, ALL_t, J j) {
  using V = View<D, P...>; using T = typename V::value_type;
  static_assert(V::rank == 2);
  return StridedColumn<T>{v.data() + (size_t)j * v.stride(1), v.extent(0), v.stride(0)};
}
template <class D, class... P, class A, class B> auto subview(const View<D, P...>& v, std::pair<A, B> r) {
  using V = View<D, P...>; using T = typename V::value_type;
  static_assert(V::rank == 1);
  return View<T*>(v.data() + (size_t)r.first, (size_t)(r.second - r.first));
}
template <class V, class... Args> using Subview = decltype(subview(std::declval<V>(), std::declval<Args>()...));

// ---------------------------------------------------------------- atomics, math, misc
namespace Impl {
template <class T, class Op> T atomic_rmw(T* p, Op op) {  // returns the old value
  if constexpr (std::is_integral_v<T> || std::is_enum_v<T>) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (!__atomic_compare_exchange_n(p, &o, op(o), true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
  } else {
    using I = std::conditional_t<sizeof(T) == 4, uint32_t, uint64_t>;
    static_assert(sizeof(T) == sizeof(I));
    I* q = reinterpret_cast<I*>(p);
    I o = __atomic_load_n(q, __ATOMIC_RELAXED);
    for (;;) {
      T ov; std::memcpy(&ov, &o, sizeof(T));
      const T nv = op(ov);
      I ni; std::memcpy(&ni, &nv, sizeof(T));
      if (__atomic_compare_exchange_n(q, &o, ni, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return ov;
    }
  }
}
}  // namespace Impl
template <class T, class U> T atomic_fetch_add(T* p, U v) { return Impl::atomic_rmw(p, [v](T o) { return (T)(o + (T)v); }); }
template <class T> T atomic_fetch_inc(T* p) { return Impl::atomic_rmw(p, [](T o) { return (T)(o + 1); }); }
template <class T, class U> void atomic_add(T* p, U v) { Impl::atomic_rmw(p, [v](T o) { return (T)(o + v); }); }
template <class T> T atomic_load(const T* p) { return Impl::atomic_rmw(const_cast<T*>(p), [](T o) { return o; }); }
template <class T, class U> T atomic_exchange(T* p, U v) { return Impl::atomic_rmw(p, [v](T) { return (T)v; }); }
template <class T, class U> void atomic_store(T* p, U v) { Impl::atomic_rmw(p, [v](T) { return (T)v; }); }
inline void fence() {}
inline void fence(const std::string&) {}
template <class... A> int printf(const char* fmt, A... a) { if constexpr (sizeof...(A) == 0) return std::fputs(fmt, stderr); else return std::fprintf(stderr, fmt, a...); }
inline void abort(const char* m) { std::fprintf(stderr, "%s\n", m); std::abort(); }

using std::abs; using std::sqrt; using std::exp; using std::pow; using std::isfinite; using std::erf; using std::erfc; using std::copysign;
using std::floor; using std::ceil; using std::tanh; using std::cbrt; using std::isnan; using std::fabs; using std::log1p; using std::expm1; using std::exp2;
using std::isinf; using std::sin; using std::cos; using std::fmin; using std::fmax; using std::round; using std::trunc; using std::log2; using std::log10;
// Kokkos::log(float) is re-specified: correctly rounded from the double logarithm
inline float log(float x) { return (float)std::log((double)x); }
inline double log(double x) { return std::log(x); }
inline long double log(long double x) { return std::log(x); }
template <class T> requires std::is_integral_v<T> double log(T x) { return std::log((double)x); }
template <class A, class B> constexpr auto max(const A& a, const B& b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }
template <class A, class B> constexpr auto min(const A& a, const B& b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <class T> constexpr T min(std::initializer_list<T> l) { return std::min(l); }
template <class T> constexpr T max(std::initializer_list<T> l) { return std::max(l); }
template <class T> constexpr const T& clamp(const T& v, const T& lo, const T& hi) { return v < lo ? lo : (hi < v ? hi : v); }

namespace numbers {
template <class T> inline constexpr T pi_v = T(3.141592653589793238462643383279502884L);
template <class T> inline constexpr T sqrt2_v = T(1.414213562373095048801688724209698079L);
template <class T> inline constexpr T ln2_v = T(0.693147180559945309417232121458176568L);
template <class T> inline constexpr T e_v = T(2.718281828459045235360287471352662498L);
template <class T> inline constexpr T inv_pi_v = T(0.318309886183790671537767526745028724L);
template <class T> inline constexpr T inv_sqrtpi_v = T(0.564189583547756286948079451560772586L);
template <class T> inline constexpr T ln10_v = T(2.302585092994045684017991454684364208L);
inline constexpr double pi = pi_v<double>;
inline constexpr double sqrt2 = sqrt2_v<double>;
inline constexpr double ln2 = ln2_v<double>;
inline constexpr double e = e_v<double>;
inline constexpr double inv_pi = inv_pi_v<double>;
inline constexpr double inv_sqrtpi = inv_sqrtpi_v<double>;
inline constexpr double ln10 = ln10_v<double>;
}  // namespace numbers

namespace Experimental { struct half_t { float v; }; }
namespace Profiling { struct ScopedRegion { explicit ScopedRegion(const std::string&) {} }; inline void pushRegion(const std::string&) {} inline void popRegion() {} }

// ---------------------------------------------------------------- policies
namespace Impl {
template <class P> struct is_space : std::bool_constant<std::is_same_v<P, Serial>> {};
template <class... P> struct PickTag { using type = void; };
template <class P0, class... P> struct PickTag<P0, P...> { using type = std::conditional_t<is_space<P0>::value, typename PickTag<P...>::type, P0>; };
}  // namespace Impl

// hooks the driver installs (see the bottom of this file)
namespace shim { void team_begin(size_t league_rank); void range_index(size_t i); void kernel_begin(const std::string& label); void kernel_end(); }

struct PerTeamValue { size_t v; };
struct PerThreadValue { size_t v; };
inline PerTeamValue PerTeam(size_t v) { return {v}; }
inline PerThreadValue PerThread(size_t v) { return {v}; }

class HostTeamMember {
  size_t lr_, ls_;
  ScratchMemorySpace scratch_;
 public:
  using execution_space = Serial;
  using scratch_memory_space = ScratchMemorySpace;
  HostTeamMember(size_t lr, size_t ls, char* sb, char* se) : lr_(lr), ls_(ls) { scratch_.cur = sb; scratch_.end = se; }
  int league_rank() const { return (int)lr_; }
  int league_size() const { return (int)ls_; }
  int team_rank() const { return 0; }
  int team_size() const { return 1; }
  void team_barrier() const {}
  const ScratchMemorySpace& team_scratch(int) const { return scratch_; }
  const ScratchMemorySpace& team_shmem() const { return scratch_; }
  const ScratchMemorySpace& thread_scratch(int) const { return scratch_; }
};
inline PerTeamValue PerTeam(const HostTeamMember&) { return {0}; }
inline PerThreadValue PerThread(const HostTeamMember&) { return {0}; }

template <class... Props> class TeamPolicy {
  size_t league_ = 0, scratch_ = 0;
 public:
  using execution_space = Serial;
  using member_type = HostTeamMember;
  using work_tag = typename Impl::PickTag<Props...>::type;
  TeamPolicy() = default;
  template <class T, class V> TeamPolicy(const Serial&, size_t league, T, V) : league_(league) {}
  template <class T> TeamPolicy(const Serial&, size_t league, T) : league_(league) {}
  template <class T, class V> TeamPolicy(size_t league, T, V) : league_(league) {}
  template <class T> TeamPolicy(size_t league, T) : league_(league) {}
  TeamPolicy& set_scratch_size(int, PerTeamValue t) { scratch_ = std::max(scratch_, t.v); return *this; }
  TeamPolicy& set_scratch_size(int, PerTeamValue t, PerThreadValue h) { scratch_ = std::max(scratch_, t.v + h.v); return *this; }
  size_t league_size() const { return league_; }
  size_t scratch_size(int = 0) const { return scratch_; }
  int team_size() const { return 1; }
};

template <class... Props> class RangePolicy {
  size_t b_ = 0, e_ = 0;
 public:
  using execution_space = Serial;
  using member_type = size_t;
  using work_tag = typename Impl::PickTag<Props...>::type;
  RangePolicy() = default;
  RangePolicy(size_t b, size_t e) : b_(b), e_(e) {}
  RangePolicy(const Serial&, size_t b, size_t e) : b_(b), e_(e) {}
  size_t begin() const { return b_; }
  size_t end() const { return e_; }
};

// MDRangePolicy<Space, Rank<2, outer, inner>>({b0, b1}, {e0, e1}): the two-dimensional form the reference's liquid solver
// uses; Iterate::Left makes the FIRST index the fastest
enum class Iterate { Default, Left, Right };
template <unsigned N, Iterate Outer = Iterate::Default, Iterate Inner = Iterate::Default> struct Rank { static constexpr unsigned rank = N; static constexpr Iterate outer = Outer; };
template <class... Props> class MDRangePolicy {
  template <class P> struct IsRank : std::false_type {};
  template <unsigned N, Iterate O, Iterate I> struct IsRank<Rank<N, O, I>> : std::true_type {};
  template <class... Q> struct PickRank { using type = Rank<2>; };
  template <class Q0, class... Q> struct PickRank<Q0, Q...> { using type = std::conditional_t<IsRank<Q0>::value, Q0, typename PickRank<Q...>::type>; };
 public:
  using rank_type = typename PickRank<Props...>::type;
  static_assert(rank_type::rank == 2, "kokkos_shim: MDRangePolicy is provided for rank 2");
  template <class B, class E> MDRangePolicy(std::initializer_list<B> b, std::initializer_list<E> e) {
    size_t k = 0; for (auto x : b) b_[k++] = (size_t)x;
    k = 0; for (auto x : e) e_[k++] = (size_t)x;
  }
  size_t b_[2] = {0, 0}, e_[2] = {0, 0};
};
template <class F, class... P> void parallel_for(const std::string& label, const MDRangePolicy<P...>& pol, const F& f) {
  shim::kernel_begin(label);
  if (MDRangePolicy<P...>::rank_type::outer == Iterate::Left) { for (size_t j = pol.b_[1]; j < pol.e_[1]; ++j) for (size_t i = pol.b_[0]; i < pol.e_[0]; ++i) f((int)i, (int)j); }
  else { for (size_t i = pol.b_[0]; i < pol.e_[0]; ++i) for (size_t j = pol.b_[1]; j < pol.e_[1]; ++j) f((int)i, (int)j); }
  shim::kernel_end();
}

struct NestedRange { size_t b, e; };
inline NestedRange TeamThreadRange(const HostTeamMember&, size_t n) { return {0, n}; }
inline NestedRange TeamThreadRange(const HostTeamMember&, size_t b, size_t e) { return {b, e}; }
inline NestedRange TeamVectorRange(const HostTeamMember&, size_t n) { return {0, n}; }
inline NestedRange TeamVectorRange(const HostTeamMember&, size_t b, size_t e) { return {b, e}; }
inline NestedRange ThreadVectorRange(const HostTeamMember&, size_t n) { return {0, n}; }
inline NestedRange ThreadVectorRange(const HostTeamMember&, size_t b, size_t e) { return {b, e}; }

template <class F> void parallel_for(const NestedRange& r, const F& f) { for (size_t i = r.b; i < r.e; ++i) f(i); }
template <class F, class T> void parallel_reduce(const NestedRange& r, const F& f, T& result) {
  T tmp{};  // the reduction identity of a sum
  for (size_t i = r.b; i < r.e; ++i) f(i, tmp);
  result = tmp;
}
template <class F> void single(PerTeamValue, const F& f) { f(); }
template <class F> void single(PerThreadValue, const F& f) { f(); }

namespace Impl {
template <class Tag, class F, class... A> void call(const F& f, A&&... a) {
  if constexpr (std::is_void_v<Tag>) f(std::forward<A>(a)...); else f(Tag{}, std::forward<A>(a)...);
}
template <class Policy, class Body> void for_each_team(const Policy& pol, const Body& body) {  // body(member, thread)
  const int nt = shim::in_parallel() ? 1 : shim::n_threads();
  const long n = (long)pol.league_size();
#pragma omp parallel num_threads(nt) if (nt > 1)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    std::vector<char> scratch(pol.scratch_size() + 64);
#pragma omp for schedule(static)
    for (long l = 0; l < n; ++l) {
      shim::team_begin((size_t)l);
      HostTeamMember m((size_t)l, (size_t)n, scratch.data(), scratch.data() + scratch.size());
      body(m, t);
    }
  }
}
template <class Body> void for_each_index(size_t b, size_t e, const Body& body) {  // body(i, thread)
  const int nt = shim::in_parallel() ? 1 : shim::n_threads();
#pragma omp parallel num_threads(nt) if (nt > 1)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
#pragma omp for schedule(static)
    for (long i = (long)b; i < (long)e; ++i) { shim::range_index((size_t)i); body((size_t)i, t); }
  }
}
template <class R> concept ReducerLike = requires(const R& r) { typename R::value_type; r.reference(); };
}  // namespace Impl

// parallel_for
template <class F, class... P> void parallel_for(const std::string& label, const TeamPolicy<P...>& pol, const F& f) {
  shim::kernel_begin(label);
  Impl::for_each_team(pol, [&](const HostTeamMember& m, int) { Impl::call<typename TeamPolicy<P...>::work_tag>(f, m); });
  shim::kernel_end();
}
template <class F, class... P> void parallel_for(const std::string& label, const RangePolicy<P...>& pol, const F& f) {
  shim::kernel_begin(label);
  Impl::for_each_index(pol.begin(), pol.end(), [&](size_t i, int) { Impl::call<typename RangePolicy<P...>::work_tag>(f, i); });
  shim::kernel_end();
}
template <class F, class I> requires std::is_integral_v<I> void parallel_for(const std::string& label, I n, const F& f) {
  parallel_for(label, RangePolicy<>(0, (size_t)n), f);
}
template <class F, class... P> void parallel_for(const TeamPolicy<P...>& pol, const F& f) { parallel_for(std::string(), pol, f); }
template <class F, class... P> void parallel_for(const RangePolicy<P...>& pol, const F& f) { parallel_for(std::string(), pol, f); }
template <class F, class I> requires std::is_integral_v<I> void parallel_for(I n, const F& f) { parallel_for(std::string(), RangePolicy<>(0, (size_t)n), f); }

// parallel_reduce: the result is a reducer object, a rank-0 View or a scalar reference
namespace Impl {
// per-thread partial values (each starts from `init`), joined in thread order afterwards
template <class Policy, class F, class V, class Join> void reduce_loop(const Policy& pol, const F& f, V& v, const V& init, const Join& join) {
  using Tag = typename Policy::work_tag;
  const int nt = std::max(1, shim::n_threads());
  std::vector<V> part((size_t)nt, init);
  if constexpr (requires { pol.league_size(); }) {
    for_each_team(pol, [&](const HostTeamMember& m, int t) { call<Tag>(f, m, part[(size_t)t]); });
  } else {
    for_each_index(pol.begin(), pol.end(), [&](size_t i, int t) { call<Tag>(f, i, part[(size_t)t]); });
  }
  v = part[0];
  for (int t = 1; t < nt; ++t) join(v, part[(size_t)t]);
}
}  // namespace Impl
template <class Policy, class F, class R> requires Impl::ReducerLike<R>
void parallel_reduce(const std::string& label, const Policy& pol, const F& f, const R& red) {
  shim::kernel_begin(label);
  typename R::value_type v, init; red.init(v); red.init(init);
  Impl::reduce_loop(pol, f, v, init, [&](typename R::value_type& d, const typename R::value_type& s2) { red.join(d, s2); });
  red.reference() = v;
  shim::kernel_end();
}
template <class Policy, class F, class D, class... P> void parallel_reduce(const std::string& label, const Policy& pol, const F& f, const View<D, P...>& res) {
  shim::kernel_begin(label);
  using V = typename View<D, P...>::non_const_value_type;
  V v{};
  Impl::reduce_loop(pol, f, v, V{}, [](V& d, const V& s2) { d += s2; });
  res() = v;
  shim::kernel_end();
}
template <class Policy, class F, class T> requires(!Impl::ReducerLike<T> && !is_view_v<T> && (requires(const Policy& p) { p.league_size(); } || requires(const Policy& p) { p.begin(); }))
void parallel_reduce(const std::string& label, const Policy& pol, const F& f, T& res) {
  shim::kernel_begin(label);
  T v{};
  Impl::reduce_loop(pol, f, v, T{}, [](T& d, const T& s2) { d += s2; });
  res = v;
  shim::kernel_end();
}
template <class F, class T, class I> requires std::is_integral_v<I> void parallel_reduce(const std::string& label, I n, const F& f, T& res) {
  parallel_reduce(label, RangePolicy<>(0, (size_t)n), f, res);
}
// without a label
template <class Policy, class F, class R> requires(requires(const Policy& p) { p.league_size(); } || requires(const Policy& p) { p.begin(); })
void parallel_reduce(const Policy& pol, const F& f, R&& res) { parallel_reduce(std::string(), pol, f, std::forward<R>(res)); }
// several scalar results: f(i, a, b, ...) over a range (serial or per-thread partial sums)
template <class F, class I, class T0, class T1, class... Ts> requires std::is_integral_v<I>
void parallel_reduce(const std::string& label, I n, const F& f, T0& r0, T1& r1, Ts&... rs) {
  shim::kernel_begin(label);
  std::tuple<T0, T1, Ts...> acc{};
  for (size_t i = 0; i < (size_t)n; ++i) { shim::range_index(i); std::apply([&](auto&... a) { f((I)i, a...); }, acc); }
  std::tie(r0, r1, rs...) = acc;
  shim::kernel_end();
}

// parallel_scan: a serial execution runs the final pass only, in ascending order
template <class F, class... P> void parallel_scan(const std::string& label, const RangePolicy<P...>& pol, const F& f) {
  shim::kernel_begin(label);
  using VT = std::size_t;
  VT update{};
  for (size_t i = pol.begin(); i < pol.end(); ++i) { shim::range_index(i); f((int)i, update, true); }
  shim::kernel_end();
}

// ---------------------------------------------------------------- ScatterView
// Duplicated per thread like Kokkos' host default (ScatterDuplicated, ScatterNonAtomic): access() hands out the
// calling thread's private copy, contribute() adds the copies into the target in thread order, reset() zeroes them.
namespace Experimental {
template <class DataT, class... Props> class ScatterView {
  using V = View<DataT, Props...>;
  using T = typename V::non_const_value_type;
  V target_;
  std::shared_ptr<std::vector<T>> dup_;
  size_t n_ = 0; int nt_ = 1;
  static int max_threads() {  // OMP_NUM_THREADS may say 1 (torchrun exports it) while shim::set_threads asks for more
#ifdef _OPENMP
    return std::max(omp_get_max_threads(), shim::n_threads());
#else
    return 1;
#endif
  }
 public:
  ScatterView() = default;
  explicit ScatterView(const V& v) : target_(v), n_(v.span()), nt_(max_threads()) { dup_ = std::make_shared<std::vector<T>>(n_ * (size_t)nt_, T{}); }
  struct Ref {
    T* p;
    void operator+=(T v) const { *p += v; }  // the right-hand side is converted to the value type first, like ScatterValue::operator+=(value_type const&)
    void operator-=(T v) const { *p -= v; }
  };
  struct Access {
    V t;
    template <class... I> Ref operator()(I... i) const { return Ref{&t(i...)}; }
  };
  Access access() const {
#ifdef _OPENMP
    const int th = omp_in_parallel() ? omp_get_thread_num() : 0;
#else
    const int th = 0;
#endif
    return Access{V(dup_->data() + (size_t)th * n_, target_.extent(0), target_.extent(1), target_.extent(2))};
  }
  void reset() {  // called between kernels (serial): also the place to follow a change of the thread count
    if (!dup_) return;
    if (max_threads() > nt_) { nt_ = max_threads(); dup_->assign(n_ * (size_t)nt_, T{}); }
    else std::fill(dup_->begin(), dup_->end(), T{});
  }
  template <class W> void shim_contribute_into(const W& dst) const {
    for (int th = 0; th < nt_; ++th) for (size_t k = 0; k < n_; ++k) dst.data()[k] += (*dup_)[(size_t)th * n_ + k];
  }
};
template <class D, class... P> ScatterView<D, P...> create_scatter_view(const View<D, P...>& v) { return ScatterView<D, P...>(v); }
template <class V, class S> void contribute(const V& dst, const S& scatter) { scatter.shim_contribute_into(dst); }
}  // namespace Experimental

// ---------------------------------------------------------------- sorting API named by ParticlesContainer::_sort (never called)
template <class V> struct BinOp1D { BinOp1D() = default; BinOp1D(int, typename V::const_value_type, typename V::const_value_type) {} };
template <class V, class Op, class... R> struct BinSort {
  template <class... A> explicit BinSort(A&&...) {}
  void create_permute_vector() {}
  template <class W> void sort(const W&) {}
};

// ---------------------------------------------------------------- random numbers
// The reference draws from Kokkos::Random_XorShift1024_Pool; DESIGN.md §4 re-specifies the streams as
// Philox4x32-10 with counter {index, step, draw_block, rank}.  The generator handed out by get_state()
// asks shim::rng() which stream serves the draw that is being made.
namespace shim {
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    const uint64_t a = 0xD2511F53ull * c[0], b = 0xCD9E8D57ull * c[2];
    const uint32_t n0 = (uint32_t)(b >> 32) ^ c[1] ^ k[0], n2 = (uint32_t)(a >> 32) ^ c[3] ^ k[1];
    c[1] = (uint32_t)b; c[3] = (uint32_t)a; c[0] = n0; c[2] = n2;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  for (int q = 0; q < 4; ++q) out[q] = c[q];
}
enum class Mode { Sequential, MoveTape, Leave };
struct RngState {
  uint32_t key[2] = {0, 0};
  uint32_t rank = 0, step = 0;
  Mode mode = Mode::Sequential;
  // Sequential: blocks base+1, base+2, ... of counter {slot, step, ., rank}, four words each (model hooks)
  uint32_t slot = 0, block = 2, buf[4] = {0, 0, 0, 0}; int have = 0;
  // MoveTape: draw k of team `league` serves particle league*per_team + k/2; even = u1 (block 0), odd = u2 (block 2)
  size_t league = 0, per_team = 0, tape = 0;
  // Leave: u3 of the particle the RangePolicy is visiting (block 1)
  size_t index = 0;
  void start_sequence(uint32_t s, uint32_t base) { mode = Mode::Sequential; slot = s; block = base; have = 0; }
  uint32_t word_of_quad(size_t s, uint32_t blk) const {
    const uint32_t ctr[4] = {(uint32_t)(s >> 2), step, blk, rank}; uint32_t o[4]; philox4x32_10(ctr, key, o); return o[s & 3];
  }
  uint32_t next32() {
    if (mode == Mode::MoveTape) {
      const size_t s = league * per_team + (tape >> 1); const bool second = tape & 1; ++tape;
      if (!second) return word_of_quad(s, 0u);
      const uint32_t ctr[4] = {(uint32_t)s, step, 2u, rank}; uint32_t o[4]; philox4x32_10(ctr, key, o); return o[0];
    }
    if (mode == Mode::Leave) return word_of_quad(index, 1u);
    if (have == 0) { ++block; const uint32_t ctr[4] = {slot, step, block, rank}; philox4x32_10(ctr, key, buf); have = 4; }
    return buf[4 - (have--)];
  }
};
RngState& rng();
}  // namespace shim

#ifdef KOKKOS_SHIM_XORSHIFT
// Timing build only (libbmc_ref_release.so): a generator of the cost class of the reference's, so that the
// CPU baseline is not charged for Philox.  xorshift1024* (S. Vigna, 2014, public domain), one state per
// thread — the design of a Kokkos random pool.  Streams are NOT comparable with the checker build.
namespace shim {
struct XsState {
  uint64_t s[16]; int p = 0; bool seeded = false;
  void seed(uint64_t x) { for (auto& w : s) { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; w = z ^ (z >> 31); } seeded = true; }
  uint64_t next() {
    const uint64_t s0 = s[p]; uint64_t s1 = s[p = (p + 1) & 15];
    s1 ^= s1 << 31; s[p] = s1 ^ s0 ^ (s1 >> 11) ^ (s0 >> 30);
    return s[p] * 1181783497276652981ull;
  }
};
inline XsState& xs() {
  thread_local XsState st;
  if (!st.seeded) {
#ifdef _OPENMP
    st.seed(0x2024ull + 0x1000ull * (uint64_t)omp_get_thread_num());
#else
    st.seed(0x2024ull);
#endif
  }
  return st;
}
}  // namespace shim
#endif

template <class Device = Serial> class Random_XorShift1024 {
 public:
  static constexpr uint64_t MAX_URAND64 = ~0ull;
#ifdef KOKKOS_SHIM_XORSHIFT
  shim::XsState* st_ = &shim::xs();
  uint64_t urand64() { return st_->next(); }
  uint32_t urand() { return (uint32_t)(st_->next() >> 32); }
#else
  uint32_t urand() { return shim::rng().next32(); }
  uint64_t urand64() { const uint64_t hi = urand(); return (hi << 32) | urand(); }
#endif
  uint64_t urand64(uint64_t range) { return urand64() % range; }
  uint64_t urand64(uint64_t lo, uint64_t hi) { return lo + urand64() % (hi - lo); }
  uint32_t urand(uint32_t range) { return urand() % range; }
  uint32_t urand(uint32_t lo, uint32_t hi) { return lo + urand() % (hi - lo); }
  int rand() { return (int)(urand() >> 1); }
  int rand(int range) { return rand() % range; }
  int rand(int lo, int hi) { return lo + rand() % (hi - lo); }
  // [0,1): top 24 bits of one word / 53 bits of two words
#ifdef KOKKOS_SHIM_OPEN_UNIFORM
  // (0,1): only for the reference's own distribution test, whose Exponential<float> case takes -ln(frand()) over 4e7
  // draws and asserts finiteness (Kokkos' frand never returns 0 in practice; a 24-bit uniform does, once in 1.7e7)
  float frand() { return ((float)(urand() >> 8) + 0.5f) * (1.0f / 16777216.0f); }
#else
  float frand() { return (float)(urand() >> 8) * (1.0f / 16777216.0f); }
#endif
  float frand(float range) { return range * frand(); }
  float frand(float lo, float hi) { return (lo == 0.f && hi == 1.f) ? frand() : lo + (hi - lo) * frand(); }
  double drand() { const uint64_t v = urand64() >> 11; return (double)v * (1.0 / 9007199254740992.0); }
  double drand(double range) { return range * drand(); }
  double drand(double lo, double hi) { return (lo == 0. && hi == 1.) ? drand() : lo + (hi - lo) * drand(); }
  double normal() {  // Marsaglia polar method on drand(), as documented for Kokkos generators
    double S = 2.0, U = 0.0;
    while (S >= 1.0) { U = 2.0 * drand() - 1.0; const double V = 2.0 * drand() - 1.0; S = U * U + V * V; }
    return U * std::sqrt(-2.0 * std::log(S) / S);
  }
  double normal(double mean, double sd = 1.0) { return mean + normal() * sd; }
};
template <class Device = Serial> class Random_XorShift1024_Pool {
 public:
  using generator_type = Random_XorShift1024<Device>;
  using device_type = Device;
  Random_XorShift1024_Pool() = default;
  Random_XorShift1024_Pool(uint64_t) {}  // implicit: `return (seed);` in mc/src/prng.cpp
  void init(uint64_t, int) {}
  generator_type get_state() const { return {}; }
  generator_type get_state(int) const { return {}; }
  void free_state(const generator_type&) const {}
};
template <class Device = Serial> using Random_XorShift64 = Random_XorShift1024<Device>;
template <class Device = Serial> using Random_XorShift64_Pool = Random_XorShift1024_Pool<Device>;
namespace Impl {
template <class ViewType, class RandomPool, int loops, int rank, class IndexType> struct fill_random_functor_begin_end {
  fill_random_functor_begin_end(ViewType, RandomPool, typename ViewType::const_value_type, typename ViewType::const_value_type) {}
  void operator()(IndexType) const {}
};
}  // namespace Impl
template <class V, class Pool, class T> void fill_random(const V& v, Pool pool, T lo, T hi) {
  auto g = pool.get_state();
  for (size_t i = 0; i < v.size(); ++i) v.data()[i] = (typename V::non_const_value_type)g.drand((double)lo, (double)hi);
}

inline void print_configuration(std::ostream& os, bool = false) { os << "kokkos_shim: serial stand-in (oracle/kokkos_shim/Kokkos_Core.hpp)\n"; }
inline void initialize() {}
inline void initialize(int&, char**) {}
inline void finalize() {}
inline bool is_initialized() { return true; }
struct ScopeGuard { template <class... A> explicit ScopeGuard(A&&...) {} };

}  // namespace Kokkos
