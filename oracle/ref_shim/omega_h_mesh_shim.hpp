// omega_h_mesh_shim.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Serial, host-only stand-ins for what the reference's search LOOPS touch besides the small-vector
// types of omega_h_shim.hpp, so that check_initial_parents / find_exit_face /
// check_model_intersection / set_new_element / compute_tolerance_from_area /
// trace_particle_through_mesh / RemoveParticleOnGeometricModelExit / search_mesh
// (src/pumipic_adjacency.tpp:72-660) compile UNMODIFIED (oracle/build_ref_primitives.py):
//   Omega_h::Write / Read / HostWrite   ref-counted arrays (a const copy still writes, like a view)
//   Omega_h::Mesh                       hands out the arrays it was given: ask_elem_verts, coords,
//                                       ask_down(dim,dim-1), ask_up(dim-1,dim), ask_verts_of(dim-1);
//                                       measure_elements_real and mark_exposed_sides return arrays
//                                       supplied by the caller (the same derived arrays the oracle
//                                       uses -- what is under test here is the reference's loop logic)
//   gather_verts / gather_vectors / gather_down / get_min
//   pumipic::ParticleStructure + parallel_for   a loop over all slots: fn(row element, slot, mask)
//   Kokkos::parallel_reduce(Min), atomic_add, Timer, Profiling; MPI_Comm_rank; RecordTime ...
// Every kernel of the reference is data-parallel over slots with no cross-slot dependence other than
// counters (atomics), so the loops run under OpenMP when the file is compiled with -fopenmp -- the
// role of the reference's Kokkos OpenMP backend -- and give the same results serially.
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <tuple>
#include <string>
#include <type_traits>
#include <vector>

#include "omega_h_shim.hpp"

#define PS_LAMBDA [=]
#define OMEGA_H_LAMBDA [=]
#define printError(...) ((void)0)

typedef int MPI_Comm;
static inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return 0; }

namespace Kokkos {
template <class T>
struct Min {
  T& ref;
  explicit Min(T& r) : ref(r) {}
};
template <class F, class T>
void parallel_reduce(int n, F f, Min<T> out) {
  T v = DBL_MAX;   // identity of the Min reducer
  for (int i = 0; i < n; ++i) f(i, v);
  out.ref = v;
}
template <class F, class T>
void parallel_reduce(int n, F f, T& total) {   // sum reduction into `total`
  T v = T();
  for (int i = 0; i < n; ++i) f(i, v);
  total = v;
}
template <class T, class U> void atomic_add(T* p, U v) {
#ifdef _OPENMP
#pragma omp atomic
#endif
  *p += v;
}
struct Timer { void reset() {} double seconds() const { return 0.0; } };
namespace Profiling {
static inline void pushRegion(const char*) {}
static inline void popRegion() {}
}  // namespace Profiling
}  // namespace Kokkos

namespace Omega_h {
typedef signed char I8;
typedef int ClassId;
// Omega_h::Write: a ref-counted array whose const copies still write (like a Kokkos view).  Raw
// pointer + owner so that an access is one load; filled in parallel like Omega_h fills it; `alias`
// wraps caller memory without copying (used by the timing entry points).
template <class T>
class Write {
  std::shared_ptr<T> own_;
  T* p_ = nullptr;
  int n_ = 0;
  void alloc(int n) {
    n_ = n;
    own_ = std::shared_ptr<T>(static_cast<T*>(std::malloc(sizeof(T) * (size_t)(n > 0 ? n : 1))), std::free);
    p_ = own_.get();
  }

 public:
  Write() {}
  Write(int n, T v, const std::string& = "") {
    alloc(n);
    T* q = p_;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < n; ++i) q[i] = v;
  }
  explicit Write(int n, const std::string& = "") { alloc(n); }
  template <class H, class = decltype(std::declval<H>().write())> Write(const H& h) { *this = h.write(); }
  static Write alias(T* p, int n) { Write w; w.p_ = p; w.n_ = n; return w; }
  int size() const { return n_; }
  T& operator[](int i) const { return p_[i]; }
  T* data() const { return p_; }
};
template <class T>
class Read {
  Write<T> w_;

 public:
  Read() {}
  Read(Write<T> w) : w_(w) {}
  int size() const { return w_.size(); }
  const T& operator[](int i) const { return w_[i]; }
};
template <class T>
class HostWrite {
  Write<T> w_;

 public:
  explicit HostWrite(Write<T> w) : w_(w) {}
  explicit HostWrite(int n) : w_(n, T()) {}
  T& operator[](int i) const { return w_[i]; }
  Write<T> write() const { return w_; }
};
typedef Read<LO> LOs;
typedef Read<Real> Reals;
typedef Read<I8> Bytes;
// a failed OMEGA_H_CHECK (RefCheckFailed) must not leave an OpenMP region as an exception: it is
// noted inside and rethrown after the loop
template <class F> void parallel_for(int n, F f, const std::string& = "") {
  int failed = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int i = 0; i < n; ++i) {
    try { f(i); } catch (const RefCheckFailed&) { failed = 1; }
  }
  if (failed) throw RefCheckFailed();
}
template <class T> T get_min(Read<T> a) {
  T m = a[0];
  const int n = a.size();
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(min : m)
#endif
  for (int i = 1; i < n; ++i) if (a[i] < m) m = a[i];
  return m;
}

struct Adj {
  LOs a2ab, ab2b;
};
struct CommStub {
  MPI_Comm get_impl() const { return 0; }
};
class Mesh {
 public:
  int dim_ = 0;
  LOs elem_verts, down, side_verts, up_off, up_vals, dual_off, dual_vals, v2e_off, v2e_vals;
  int nverts_ = 0;
  std::shared_ptr<std::map<std::string, Reals>> tags = std::make_shared<std::map<std::string, Reals>>();
  Reals coords_, measure;
  Bytes exposed;
  CommStub comm_;
  int dim() const { return dim_; }
  int nelems_ = -1;
  int nelems() const { return nelems_ >= 0 ? nelems_ : measure.size(); }
  LOs ask_elem_verts() const { return elem_verts; }
  Reals coords() const { return coords_; }
  int nverts() const { return nverts_; }
  Adj ask_down(int, int low) const { Adj a; a.ab2b = low == 0 ? elem_verts : down; return a; }
  Adj get_adj(int from, int to) const { return ask_down(from, to); }
  int nents_[4] = {-1, -1, -1, -1};
  int nents(int d) const { return nents_[d]; }
  std::shared_ptr<std::map<int, Adj>> ups = std::make_shared<std::map<int, Adj>>();   // ask_up(low, dim) per low
  Adj ask_up(int low, int) const {
    Adj a;
    if (ups->count(low)) return ups->at(low);
    if (low == 0) { a.a2ab = v2e_off; a.ab2b = v2e_vals; } else { a.a2ab = up_off; a.ab2b = up_vals; }
    return a;
  }
  void set_tag(int, const std::string& name, Reals v) { (*tags)[name] = v; }
  LOs class_id;
  template <class T> Read<T> get_array(int, const std::string&) const { return class_id; }   // "class_id" only
  LOs ask_verts_of(int) const { return side_verts; }
  Adj ask_dual() const { Adj a; a.a2ab = dual_off; a.ab2b = dual_vals; return a; }
  const CommStub* comm() const { return &comm_; }
};
template <int n> struct BBox { Vector<n> min, max; };
template <int n> BBox<n> get_bounding_box(Mesh* m) {
  BBox<n> b;
  Reals c = m->coords();
  for (int j = 0; j < n; ++j) b.min[j] = b.max[j] = c[j];
  for (int i = 1; i < c.size() / n; ++i)
    for (int j = 0; j < n; ++j) {
      if (c[i * n + j] < b.min[j]) b.min[j] = c[i * n + j];
      if (c[i * n + j] > b.max[j]) b.max[j] = c[i * n + j];
    }
  return b;
}
static inline Reals measure_elements_real(Mesh* m) { return m->measure; }
static inline Bytes mark_exposed_sides(Mesh* m) { return m->exposed; }

template <int n> Few<LO, n> gather_verts(LOs const& a, LO e) {
  Few<LO, n> v;
  for (int i = 0; i < n; ++i) v[i] = a[e * n + i];
  return v;
}
template <int n> Few<LO, n> gather_down(LOs const& a, LO e) { return gather_verts<n>(a, e); }
template <int neev, int dim> Matrix<dim, neev> gather_vectors(Reals const& a, Few<LO, neev> v) {
  Matrix<dim, neev> x;
  for (int i = 0; i < neev; ++i)
    for (int j = 0; j < dim; ++j) x[i][j] = a[v[i] * dim + j];
  return x;
}
}  // namespace Omega_h

static inline double pumipic_prebarrier(MPI_Comm) { return 0.0; }
static inline void RecordTime(const std::string&, double, double = 0.0) {}
static inline void PrintAdditionalTimeInfo(const char*, int) {}

namespace pumipic {
using ::RecordTime;
typedef int lid_t;
typedef double fp_t;
typedef fp_t Vector3d[3];

template <class T> struct MemberBase { typedef T type; static constexpr int ncomp = 1; };
template <class T, int N> struct MemberBase<T[N]> { typedef T type; static constexpr int ncomp = N; };
template <class... Ts> struct MemberTypes {
  static constexpr std::size_t size = sizeof...(Ts);
  template <std::size_t N> using type = typename std::tuple_element<N, std::tuple<Ts...>>::type;
};
// value(particle, component) at base[component * stride + particle]; a const copy still writes
template <class T> struct MemberAccessor {
  typedef typename MemberBase<T>::type Base;
  Base* p = nullptr;
  long stride = 0;
  Base& operator()(int i) const { return p[i]; }
  Base& operator()(int i, int c) const { return p[(long)c * stride + i]; }
};
struct MemberViews {
  std::vector<void*> arrays;
  long n = 0;
};
typedef MemberViews* MemberTypeViews;
template <class DT, std::size_t... I> void alloc_members(MemberViews* v, long n, std::index_sequence<I...>) {
  (void)std::initializer_list<int>{(v->arrays.push_back(std::calloc(
      (size_t)(n > 0 ? n : 1) * MemberBase<typename DT::template type<I>>::ncomp,
      sizeof(typename MemberBase<typename DT::template type<I>>::type))), 0)...};
}
template <class DT> MemberTypeViews createMemberViews(int n) {
  MemberViews* v = new MemberViews();
  v->n = n;
  alloc_members<DT>(v, n, std::make_index_sequence<DT::size>());
  return v;
}
template <class DT, std::size_t N> MemberAccessor<typename DT::template type<N>> getMemberView(MemberTypeViews v) {
  MemberAccessor<typename DT::template type<N>> a;
  a.p = static_cast<typename MemberBase<typename DT::template type<N>>::type*>(v->arrays[N]);
  a.stride = v->n;
  return a;
}
template <class DT> void destroyViews(MemberTypeViews v) {
  for (void* p : v->arrays) std::free(p);
  delete v;
}

template <class T> class KView {   // Kokkos::View<T*>: ("name", n), (i) indexing, shared storage
  std::shared_ptr<std::vector<T>> d_;

 public:
  KView() : d_(std::make_shared<std::vector<T>>()) {}
  KView(const std::string&, int n) : d_(std::make_shared<std::vector<T>>((size_t)n, T())) {}
  T& operator()(int i) const { return (*d_)[(size_t)i]; }
  T& operator[](int i) const { return (*d_)[(size_t)i]; }
  int size() const { return (int)d_->size(); }
};

   // particle_structs/src/support/ppTypes.h
// what ps::parallel_for needs from a structure: capacity, row element and mask per slot
template <class DataTypes>
class ParticleStructure {
 public:
  int cap = 0;
  const int* slot_elem = nullptr;
  const unsigned char* mask = nullptr;
  typedef KView<int> kkLidView;
  typedef KView<long> kkGidView;
  MemberViews* members = nullptr;
  virtual ~ParticleStructure() {}
  int capacity() const { return cap; }
  template <std::size_t N> MemberAccessor<typename DataTypes::template type<N>> get() {
    return getMemberView<DataTypes, N>(members);
  }
};
template <class DataTypes, class F>
void parallel_for(ParticleStructure<DataTypes>* ps, F& fn, std::string = "") {
  const int cap = ps->cap;
  const int* se = ps->slot_elem;
  const unsigned char* mk = ps->mask;
  int failed = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int s = 0; s < cap; ++s) {
    try { fn(se[s], s, mk[s] != 0); } catch (const RefCheckFailed&) { failed = 1; }
  }
  if (failed) throw RefCheckFailed();
}
}  // namespace pumipic
namespace particle_structs = pumipic;
