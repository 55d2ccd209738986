// ref_xgcm_init.cpp -- TEST INFRASTRUCTURE ONLY.  The generator of pseudoXGCm's particle load
// (test/pseudoXGCm.cpp: setSourceElements -- a normal number of particles per owned element whose
// class id is at most mdlFace, std::default_random_engine(1024*1024) -- and setInitialPtclCoords --
// a uniform point in the row's triangle per SLOT, std::default_random_engine(512*512)), extracted
// into ref_xgcm_init.inc (a build-time temporary) and compiled unmodified: what
// pumi-pic_b200/workloads.py xgc_source_elements / xgc_initial_coords restate.
#include <random>

#include "xgcm_shim.hpp"

namespace o = Omega_h;
namespace p = pumipic;
namespace ps = particle_structs;
using particle_structs::lid_t;
using particle_structs::MemberTypes;
using pumipic::fp_t;
using pumipic::Vector3d;

namespace xgcm_init {
#include "ref_xgcm_init.inc"
}  // namespace xgcm_init

namespace {
template <class T> o::Write<T> to_w(const T* a, long n) {
  o::Write<T> w((int)n, T());
  for (long i = 0; i < n; ++i) w[(int)i] = a[i];
  return w;
}
}  // namespace

extern "C" {
// returns the particle total; ppe_out [nelems]
int ref_xgcm_source_elements(int nelems, const int* class_id, const int* owners, int self, int mdl_face,
                             int num_ptcls, int* ppe_out) {
  p::Mesh pic;
  pic.omesh.dim_ = 2;
  pic.omesh.nelems_ = nelems;
  pic.omesh.class_id = o::LOs(to_w(class_id, nelems));
  pic.owners = o::LOs(to_w(owners, nelems));
  pic.comm_.rank_ = self;
  xgcm_init::PS::kkLidView ppe("ppe", nelems);
  const int np = xgcm_init::setSourceElements(pic, ppe, mdl_face, num_ptcls);
  for (int i = 0; i < nelems; ++i) ppe_out[i] = ppe(i);
  return np;
}
// x: [3][stride] doubles, written on masked slots
void ref_xgcm_initial_coords(int nverts, const double* coords, int nelems, const int* elem2verts, int cap,
                             const int* slot_elem, const unsigned char* mask, long stride, double* x) {
  p::Mesh pic;
  pic.omesh.dim_ = 2;
  pic.omesh.nverts_ = nverts;
  pic.omesh.nelems_ = nelems;
  pic.omesh.coords_ = o::Reals(to_w(coords, 2L * nverts));
  pic.omesh.elem_verts = o::LOs(to_w(elem2verts, 3L * nelems));
  xgcm_init::PS ptcls;
  pumipic::MemberViews mv;
  mv.arrays = {x, nullptr, nullptr, nullptr, nullptr};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
  xgcm_init::setInitialPtclCoords(pic, &ptcls, false);
}
}
