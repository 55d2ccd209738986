// ref_testadj.cpp -- TEST INFRASTRUCTURE ONLY.  The generator of the reference's synthetic search
// workload (test/test_adj.cpp: setSourceElements, init2DInternal, init3DInternal,
// get_push_distance, push_ptcls), extracted into ref_testadj.inc (a build-time temporary) and compiled unmodified:
// what pumi-pic_b200/workloads.py restates (std::default_random_engine(512*512) drawing per slot,
// the fold into the simplex, the direction on the sphere, the push distance).
#include <random>

#include "xgcm_shim.hpp"

namespace o = Omega_h;
namespace p = pumipic;
namespace ps = particle_structs;
using p::fp_t;
using p::Vector3d;
#define TriVerts 3
#define TriDim 2

namespace pumipic {
#include "ref_primitives.inc"
}  // namespace pumipic

#include "ref_testadj.inc"

namespace {
template <class T> o::Write<T> to_w(const T* a, long n) {
  o::Write<T> w((int)n, T());
  for (long i = 0; i < n; ++i) w[(int)i] = a[i];
  return w;
}
o::Mesh make_mesh(int dim, int nverts, const double* coords, int nelems, const int* elem2verts) {
  o::Mesh m;
  m.dim_ = dim; m.nverts_ = nverts; m.nelems_ = nelems;
  m.coords_ = o::Reals(to_w(coords, (long)dim * nverts));
  m.elem_verts = o::LOs(to_w(elem2verts, (long)(dim + 1) * nelems));
  return m;
}
}  // namespace

extern "C" {
int ref_testadj_ppe(int nelems, int num_ptcls, int* ppe_out) {
  o::Mesh m;
  m.nelems_ = nelems; m.dim_ = 3;
  PS::kkLidView ppe("ppe", nelems);
  const int tot = setSourceElements(m, ppe, num_ptcls);
  for (int i = 0; i < nelems; ++i) ppe_out[i] = ppe(i);
  return tot;
}
double ref_testadj_push_distance(int dim, int nverts, const double* coords, int nelems) {
  o::Mesh m;
  m.dim_ = dim; m.nverts_ = nverts; m.nelems_ = nelems;
  m.coords_ = o::Reals(to_w(coords, (long)dim * nverts));
  return get_push_distance(m);
}
// x, xtgt, motion: [3][stride] doubles; pids [cap]
void ref_testadj_init_internal(int dim, int nverts, const double* coords, int nelems, const int* elem2verts,
                               const double* measure, int cap, const int* slot_elem, const unsigned char* mask, long stride, double* x, double* xtgt,
                               int* pids, double* motion) {
  o::Mesh m = make_mesh(dim, nverts, coords, nelems, elem2verts);
  m.measure = o::Reals(to_w(measure, (long)nelems));   // measure_elements_real (init2DInternal's self-check)
  PS ptcls; pumipic::MemberViews mv;
  mv.arrays = {x, xtgt, pids, motion};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
  if (dim == 2) init2DInternal(m, &ptcls); else init3DInternal(m, &ptcls);
}
void ref_testadj_push(int cap, const int* slot_elem, const unsigned char* mask, long stride, double* x, double* xtgt,
                      int* pids, double* motion, double distance) {
  PS ptcls; pumipic::MemberViews mv;
  mv.arrays = {x, xtgt, pids, motion};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
  push_ptcls(&ptcls, distance);
}

// ---- timing entry points for bench.py's CPU legs: one step = the reference's push_ptcls
// (test_adj.cpp:550-562) followed by the reference's search_mesh (adjacency.tpp:642), both compiled
// unmodified, OpenMP loops.  The mesh arrays are copied once; Omega_h's per-call derivations
// (measure_elements_real, mark_exposed_sides x2, ask_up: adjacency.tpp:489-501,621) are NOT redone
// per step here -- the stand-in hands back precomputed arrays -- which favours the reference.
struct RefBench {
  o::Mesh mesh;
  PS ptcls;
  pumipic::MemberViews mv;
  std::vector<int> slot_elem, pids;
  std::vector<unsigned char> mask;
};
void* ref_bench_create(int dim, int nverts, const double* coords, int nelems, const int* elem2verts, int nsides,
                       const int* elem2sides, const int* side2verts, const int* side2elem_off, const int* side2elem,
                       const signed char* exposed, const double* measure, int cap, const int* slot_elem,
                       const unsigned char* mask) {
  RefBench* b = new RefBench();
  b->mesh = make_mesh(dim, nverts, coords, nelems, elem2verts);
  b->mesh.down = o::LOs(to_w(elem2sides, (long)(dim + 1) * nelems));
  b->mesh.side_verts = o::LOs(to_w(side2verts, (long)dim * nsides));
  b->mesh.up_off = o::LOs(to_w(side2elem_off, (long)nsides + 1));
  b->mesh.up_vals = o::LOs(to_w(side2elem, (long)side2elem_off[nsides]));
  b->mesh.exposed = o::Bytes(to_w(exposed, (long)nsides));
  b->mesh.measure = o::Reals(to_w(measure, (long)nelems));
  b->slot_elem.assign(slot_elem, slot_elem + cap);
  b->mask.assign(mask, mask + cap);
  b->pids.resize((size_t)cap);
  for (int i = 0; i < cap; ++i) b->pids[(size_t)i] = i;
  b->ptcls.cap = cap; b->ptcls.slot_elem = b->slot_elem.data(); b->ptcls.mask = b->mask.data();
  b->ptcls.members = &b->mv;
  return b;
}
void ref_bench_destroy(void* h) { delete static_cast<RefBench*>(h); }
// x: origins [3][stride] (in), xtgt: [3][stride] (out: x + distance*dir on masked slots), dir [3][stride];
// elem_ids [cap]: in (carried over from the previous step) unless ids_empty, out.  Returns found.
int ref_bench_step(void* h, double* x, double* xtgt, double* dir, long stride, double distance, int* elem_ids,
                   int ids_empty) {
  RefBench* b = static_cast<RefBench*>(h);
  const int cap = b->ptcls.cap;
  b->mv.arrays = {x, xtgt, b->pids.data(), dir};
  b->mv.n = stride;
  const long n3 = 3 * stride;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < n3; ++i) xtgt[i] = x[i];          // tgt starts at the current position (test_adj.cpp:503)
  push_ptcls(&b->ptcls, distance);
  auto cur = b->ptcls.get<0>();
  auto tgt = b->ptcls.get<1>();
  auto pids = b->ptcls.get<2>();
  o::Write<o::LO> ids, faces;
  o::Write<o::Real> pts;
  if (!ids_empty) ids = o::Write<o::LO>::alias(elem_ids, cap);
  const bool found = pumipic::search_mesh(b->mesh, &b->ptcls, cur, tgt, pids, ids, false, faces, pts, 0, 0);
  if (ids_empty) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < cap; ++i) elem_ids[i] = ids[i];
  }
  return found;
}
}
