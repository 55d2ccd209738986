// ref_xgcm.cpp -- TEST INFRASTRUCTURE ONLY.  The reference's gyro ring mapping and gyro scatter
// (test/gyroScatter.hpp, extracted into ref_xgcm.inc (a build-time temporary) by build_ref_primitives.py)
// compiled unmodified against xgcm_shim.hpp, behind extern "C" entry points with the oracle's
// argument lists.
#include "xgcm_shim.hpp"

namespace o = Omega_h;
namespace p = pumipic;
namespace ps = particle_structs;
using particle_structs::lid_t;
using particle_structs::MemberTypes;
using particle_structs::SellCSigma;
using pumipic::fp_t;
using pumipic::Vector3d;
#define TriVerts 3   /* src/pumipic_adjacency.hpp:68-69 */
#define TriDim 2

namespace pumipic {
#include "ref_primitives.inc"
}  // namespace pumipic

#include "ref_xgcm.inc"

namespace pumipic {
#include "ref_ptcl_ops.inc"
}  // namespace pumipic

namespace {
template <class T> o::Write<T> to_w(const T* a, long n) {
  o::Write<T> w((int)n, T());
  for (long i = 0; i < n; ++i) w[(int)i] = a[i];
  return w;
}
void fill(o::Mesh& mesh, int nverts, const double* coords, int nelems, const int* elem2verts, int nsides,
          const int* elem2sides, const int* side2verts, const int* side2elem_off, const int* side2elem,
          const signed char* exposed, const double* measure, const int* v2e_off, const int* v2e) {
  mesh.dim_ = 2;
  mesh.nverts_ = nverts;
  mesh.coords_ = o::Reals(to_w(coords, 2L * nverts));
  mesh.elem_verts = o::LOs(to_w(elem2verts, 3L * nelems));
  mesh.down = o::LOs(to_w(elem2sides, 3L * nelems));
  mesh.side_verts = o::LOs(to_w(side2verts, 2L * nsides));
  mesh.up_off = o::LOs(to_w(side2elem_off, (long)nsides + 1));
  mesh.up_vals = o::LOs(to_w(side2elem, (long)side2elem_off[nsides]));
  mesh.exposed = o::Bytes(to_w(exposed, (long)nsides));
  mesh.measure = o::Reals(to_w(measure, (long)nelems));
  mesh.v2e_off = o::LOs(to_w(v2e_off, (long)nverts + 1));
  mesh.v2e_vals = o::LOs(to_w(v2e, (long)v2e_off[nverts]));
}
}  // namespace

extern "C" {
// createGyroRingMappings (test/gyroScatter.hpp:96-166): forward map out [3 * nverts * nrings * ppr]
// (the backward map is built from the same points and is identical, :127-131)
int ref_gyro_ring_map(int nverts, const double* coords, int nelems, const int* elem2verts, int nsides,
                      const int* elem2sides, const int* side2verts, const int* side2elem_off,
                      const int* side2elem, const signed char* exposed, const double* measure,
                      const int* v2e_off, const int* v2e, double rmax, int nrings, int ppr, int theta_deg,
                      int* fwd_map, int* maps_equal) {
  o::Mesh mesh;
  fill(mesh, nverts, coords, nelems, elem2verts, nsides, elem2sides, side2verts, side2elem_off, side2elem, exposed,
       measure, v2e_off, v2e);
  setGyroConfig(rmax, nrings, ppr, theta_deg);
  o::LOs fwd, bkwd;
  createGyroRingMappings(&mesh, fwd, bkwd);
  int same = fwd.size() == bkwd.size();
  for (int i = 0; i < fwd.size(); ++i) { fwd_map[i] = fwd[i]; same = same && fwd[i] == bkwd[i]; }
  *maps_equal = same;
  return fwd.size();
}

// gyroScatter (test/gyroScatter.hpp:168-229): vertex weights out [nverts]
void ref_gyro_scatter(int nverts, int nelems, const int* elem2verts, int cap, const int* slot_elem,
                      const unsigned char* mask, const int* v2v, long v2v_len, double rmax, int nrings, int ppr,
                      double* scatter_w) {
  o::Mesh mesh;
  mesh.dim_ = 2;
  mesh.nverts_ = nverts;
  mesh.elem_verts = o::LOs(to_w(elem2verts, 3L * nelems));
  setGyroConfig(rmax, nrings, ppr, 0);
  PS ptcls;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask;
  gyroScatter(&mesh, &ptcls, o::LOs(to_w(v2v, v2v_len)), "w");
  o::Reals w = (*mesh.tags)["w"];
  for (int v = 0; v < nverts; ++v) scatter_w[v] = w[v];
}

// ellipticalPush::setup / ellipticalPush::push (test/ellipticalPush.hpp:10-70) on a structure whose
// member arrays are the caller's: x, xtgt [3][stride] doubles, b and phi floats
static void wrap_particles(PS& ptcls, pumipic::MemberViews& mv, int cap, const int* slot_elem,
                           const unsigned char* mask, double* x, double* xtgt, long stride, float* b, float* phi,
                           std::vector<int>& ids) {
  ids.assign((size_t)cap, 0);
  mv.arrays = {x, xtgt, ids.data(), b, phi};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
}
void ref_elliptical_setup(int cap, const int* slot_elem, const unsigned char* mask, double* x, long stride,
                          float* b, float* phi, double h, double k, double d) {
  PS ptcls; pumipic::MemberViews mv; std::vector<int> ids;
  wrap_particles(ptcls, mv, cap, slot_elem, mask, x, x, stride, b, phi, ids);
  ellipticalPush::setup(&ptcls, h, k, d);
}
void ref_elliptical_push(int cap, const int* slot_elem, const unsigned char* mask, double* xtgt, long stride,
                         float* b, float* phi, int nelems, const int* class_ids, double h, double k, double d,
                         double deg) {
  PS ptcls; pumipic::MemberViews mv; std::vector<int> ids;
  wrap_particles(ptcls, mv, cap, slot_elem, mask, xtgt, xtgt, stride, b, phi, ids);
  o::Mesh mesh;
  mesh.dim_ = 2;
  mesh.class_id = o::LOs(to_w(class_ids, (long)nelems));
  ellipticalPush::h = h; ellipticalPush::k = k; ellipticalPush::d = d;
  ellipticalPush::push(&ptcls, mesh, deg, 0);
}

// setUnsafeProcs (src/pumipic_ptcl_ops.hpp:33-53)
void ref_set_unsafe_procs(int cap, const int* slot_elem, const unsigned char* mask, const int* elems, int nelems,
                          const int* safe, const int* owner, int self, int* new_elems, int* new_procs) {
  pumipic::Mesh mesh;
  mesh.omesh.dim_ = 3;
  mesh.owners = o::LOs(to_w(owner, (long)nelems));
  mesh.safe = o::LOs(to_w(safe, (long)nelems));
  mesh.comm_.rank_ = self;
  PS ptcls;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask;
  PS::kkLidView ne("ne", cap), np("np", cap);
  pumipic::setUnsafeProcs(mesh, &ptcls, o::LOs(to_w(elems, (long)cap)), ne, np);
  for (int i = 0; i < cap; ++i) { new_elems[i] = ne(i); new_procs[i] = np(i); }
}
}
