// xgcm_shim.hpp -- TEST INFRASTRUCTURE ONLY.
//
// What the reference's gyro scatter (test/gyroScatter.hpp: setGyroConfig, searchAndBuildMap,
// createGyroRingMappings, gyroScatter) needs on top of omega_h_mesh_shim.hpp to compile UNMODIFIED:
// member-typed particle structures (MemberTypes, createMemberViews / getMemberView, a SellCSigma
// whose slot i simply is particle i -- results are keyed by particle id, so the layout is free),
// Kokkos views with (i) indexing, atomic_fetch_add, cos / sin, Omega_h's average().
#pragma once
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

#include "omega_h_mesh_shim.hpp"

#define MPI_COMM_WORLD 0

namespace Kokkos {
template <class T, class U> T atomic_fetch_add(T* p, U v) {
  T old;
#ifdef _OPENMP
#pragma omp atomic capture
#endif
  { old = *p; *p += v; }
  return old;
}
static inline double round(double x) { return std::round(x); }
struct DefaultExecutionSpace {};
template <class Space> struct TeamPolicy { int league, team; };
}  // namespace Kokkos

namespace Omega_h {
template <class T> T get_sum(Read<T> a) {
  T s = T();
  for (int i = 0; i < a.size(); ++i) s += a[i];
  return s;
}
template <int dim, int n> Vector<dim> average(Matrix<dim, n> x) {
  Vector<dim> avg = x[0];
  for (int i = 1; i < n; ++i) avg = avg + x[i];
  return avg / n;
}
}  // namespace Omega_h

namespace pumipic {
// slot i = particle i (any layout is legal: every consumer keys its results by particle id)
template <class DataTypes> class SellCSigma : public ParticleStructure<DataTypes> {
  std::vector<int> elems_;
  std::vector<unsigned char> mask_;

 public:
  template <class Policy>
  SellCSigma(Policy&, int, int, int, int np, KView<int>, KView<long>, KView<int> particle_elements,
             MemberTypeViews particle_info) {
    elems_.resize((size_t)np);
    mask_.assign((size_t)np, 1);
    for (int i = 0; i < np; ++i) elems_[(size_t)i] = particle_elements(i);
    this->cap = np; this->slot_elem = elems_.data(); this->mask = mask_.data();
    this->members = particle_info;   // the reference copies; sharing is equivalent here (read back by id)
  }
};
// pumipic::Mesh (src/pumipic_mesh.hpp) as far as setUnsafeProcs reads it
class Mesh {
 public:
  struct Comm { int rank_ = 0; int rank() const { return rank_; } };
  Omega_h::Mesh omesh;
  Omega_h::LOs owners, safe;
  Comm comm_;
  Omega_h::Mesh* operator->() { return &omesh; }
  Omega_h::Mesh* mesh() { return &omesh; }
  int dim() const { return omesh.dim(); }
  Omega_h::LOs entOwners(int) const { return owners; }
  Omega_h::LOs safeTag() const { return safe; }
  const Comm* comm() const { return &comm_; }
};
static inline Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace> TeamPolicyAuto(int league, int team) {
  return Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace>{league, team};
}
}  // namespace pumipic
