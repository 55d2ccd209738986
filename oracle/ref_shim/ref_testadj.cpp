// ref_testadj.cpp -- TEST INFRASTRUCTURE ONLY.  The generator of the reference's synthetic search
// workload (test/test_adj.cpp: setSourceElements, init2DInternal, init3DInternal,
// get_push_distance, push_ptcls), extracted into oracle/_ref/ref_testadj.inc and compiled unmodified:
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
}
