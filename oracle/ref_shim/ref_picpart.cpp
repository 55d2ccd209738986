// ref_picpart.cpp -- TEST INFRASTRUCTURE ONLY.  The set-up kernels of PICpart construction
// (src/pumipic_part_construct.cpp: setOwnerByClassification, BFS, bfsBufferLayers, bfsSafeInward, defineOwners,
// calculateOwnerOffset, GlobalNumberer / createGlobalNumbering, rankLidNumbering), extracted into
// ref_picpart.inc (a build-time temporary) and compiled unmodified: what
// pumi-pic_b200/csrc/pp_host_picpart.cpp (picpart_tags) and pp_host_ppm.cpp (build_world) restate.
// The few lines of Mesh::Mesh(Input&) that call them (:80-108) are a constructor body and cannot be
// lifted out; ref_picpart_tags below follows them call for call.
#include <memory>

#include "omega_h_mesh_shim.hpp"

namespace Omega_h {
typedef long long GO;
typedef Read<GO> GOs;
struct Comm {
  int rank_ = 0, size_ = 1;
  int rank() const { return rank_; }
  int size() const { return size_; }
};
typedef std::shared_ptr<Comm> CommPtr;
inline LOs offset_scan(LOs a) {
  Write<LO> out(a.size() + 1, 0);
  for (int i = 0; i < a.size(); ++i) out[i + 1] = out[i] + a[i];
  return LOs(out);
}
// Kokkos-style scan functor with an array value: serial, final pass only (same result as the
// two-pass parallel scan, which is exact for integer counts)
template <class F> void parallel_scan(int n, F f) {
  std::vector<LO> vals(f.value_count);
  f.init(vals.data());
  for (int i = 0; i < n; ++i) f((typename F::size_type)i, vals.data(), true);
}
}  // namespace Omega_h
#undef OMEGA_H_DEVICE
#define OMEGA_H_DEVICE inline
namespace Kokkos {
template <class T, class U> T atomic_fetch_add(T* p, U v) {
  T old;
#ifdef _OPENMP
#pragma omp atomic capture
#endif
  { old = *p; *p += v; }
  return old;
}
}  // namespace Kokkos

#undef printInfo
#undef printError
#include <cassert>
namespace pumipic {
inline void printInfo(const char*, ...) {}
inline void printError(const char*, ...) {}
}  // namespace pumipic

namespace picpart_ref {
#include "ref_picpart.inc"
}  // namespace picpart_ref

namespace {
namespace o = Omega_h;
template <class T> o::Write<T> to_w(const T* a, long n) {
  o::Write<T> w((int)n, T());
  for (long i = 0; i < n; ++i) w[(int)i] = a[i];
  return w;
}
o::Adj up_of(int nents, const int* off, const int* val) {
  o::Adj a;
  a.a2ab = o::LOs(to_w(off, (long)nents + 1));
  a.ab2b = o::LOs(to_w(val, (long)off[nents]));
  return a;
}
}  // namespace

extern "C" {
// Mesh::Mesh(Input&) :80-108 around the extracted bfsBufferLayers / bfsSafeInward.
// methods: 0 FULL, 1 BFS, 2 MINIMUM, 3 NONE (pumipic_input.hpp:33-39); the Input constructor's
// NONE -> MINIMUM of the buffer method and the MINIMUM -> 0 layers are applied by the caller.
void ref_picpart_tags(int dim, int bridge_dim, int nbridges, const int* up_off, const int* up_val, int nelems,
                      const int* owner, int nranks, int rank, int buffer_method, int safe_method, int buffer_layers,
                      int safe_layers, int* safe_out, int* has_part_out) {
  enum { FULL = 0, BFS = 1, MINIMUM = 2, NONE = 3 };
  o::Mesh mesh;
  mesh.dim_ = dim;
  mesh.nelems_ = nelems;
  mesh.nents_[bridge_dim] = nbridges;
  (*mesh.ups)[bridge_dim] = up_of(nbridges, up_off, up_val);
  o::CommPtr comm = std::make_shared<o::Comm>();
  comm->rank_ = rank; comm->size_ = nranks;
  o::LOs owners(to_w(owner, nelems));
  o::Write<o::LO> is_safe(nelems, safe_method == FULL);
  o::Write<o::LO> has_part(nranks, 1);
  if ((safe_method != NONE && safe_method != FULL) || buffer_method != FULL) {
    o::Write<o::LO> safe(nelems, 0);
    o::Write<o::LO> part(nranks, 0);
    picpart_ref::bfsBufferLayers(mesh, bridge_dim, comm, safe_layers, buffer_layers, safe, owners, part);
    if (safe_method == BFS || safe_method == MINIMUM) is_safe = safe;
    if (buffer_method == BFS || buffer_method == MINIMUM) has_part = part;
  }
  if (buffer_method == BFS && safe_method == FULL)
    picpart_ref::bfsSafeInward(mesh, bridge_dim, comm, safe_layers, owners, o::LOs(has_part), is_safe);
  for (int e = 0; e < nelems; ++e) safe_out[e] = is_safe[e];
  for (int p = 0; p < nranks; ++p) has_part_out[p] = has_part[p];
}

// setOwnerByClassification :278-301: element owners through the class ids (`.cpn` partitions)
void ref_owner_by_classification(int dim, int nelems, const int* class_id, int ntable, const int* class_owners,
                                 int self, int* owns_out) {
  o::Mesh mesh;
  mesh.dim_ = dim;
  mesh.nelems_ = nelems;
  mesh.class_id = o::LOs(to_w(class_id, nelems));
  o::Write<o::LO> owns(nelems, -1);
  picpart_ref::setOwnerByClassification(mesh, o::LOs(to_w(class_owners, ntable)), self, owns);
  for (int e = 0; e < nelems; ++e) owns_out[e] = owns[e];
}

// defineOwners :304-323 for the entities of one dimension
void ref_define_owners(int dim, int ent_dim, int nents, const int* up_off, const int* up_val, int nelems,
                       const int* elem_owner, int nranks, int* ent_owner_out) {
  o::Mesh mesh;
  mesh.dim_ = dim;
  mesh.nelems_ = nelems;
  mesh.nents_[ent_dim] = nents;
  (*mesh.ups)[ent_dim] = up_of(nents, up_off, up_val);
  o::CommPtr comm = std::make_shared<o::Comm>();
  comm->size_ = nranks;
  o::LOs out = picpart_ref::defineOwners(mesh, ent_dim, comm, o::LOs(to_w(elem_owner, nelems)));
  for (int i = 0; i < nents; ++i) ent_owner_out[i] = out[i];
}

// createGlobalNumbering :366-374 + rankLidNumbering :376-385: offsets [nranks+1], gids, rank-local ids
void ref_global_numbering(int nents, const int* owner, int nranks, int* offsets_out, long long* gids_out,
                          int* rank_lids_out) {
  o::LOs owners(to_w(owner, nents));
  o::Write<o::GO> gids(nents, 0);
  o::LOs off = picpart_ref::createGlobalNumbering(owners, nranks, gids);
  o::LOs lids = picpart_ref::rankLidNumbering(owners, off, o::GOs(gids));
  for (int p = 0; p <= nranks; ++p) offsets_out[p] = off[p];
  for (int i = 0; i < nents; ++i) { gids_out[i] = gids[i]; rank_lids_out[i] = lids[i]; }
}
}
