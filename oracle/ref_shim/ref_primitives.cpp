// ref_primitives.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" entry points, with the same
// signatures as the oracle's primitives (oracle/pumipic_oracle.h:60-83), around the reference's own
// functions: their text is pulled in from ref_primitives.inc (a build-time temporary), which
// oracle/build_ref_primitives.py extracts from /root/reference at build time.
#include "omega_h_mesh_shim.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace o = Omega_h;
namespace ps = particle_structs;
#define TriVerts 3   /* src/pumipic_adjacency.hpp:68-69 */
#define TriDim 2

namespace pumipic {
#include "ref_primitives.inc"
}  // namespace pumipic

namespace {
o::Vector<3> v3(const double* p) { o::Vector<3> v; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; return v; }
o::Vector<2> v2(const double* p) { o::Vector<2> v; v[0] = p[0]; v[1] = p[1]; return v; }
o::Matrix<3, 4> tet(const double M[12]) { o::Matrix<3, 4> m; for (int i = 0; i < 4; ++i) m[i] = v3(M + 3 * i); return m; }
o::Few<o::Vector<3>, 3> tri3(const double f[9]) { o::Few<o::Vector<3>, 3> t; for (int i = 0; i < 3; ++i) t[i] = v3(f + 3 * i); return t; }
}  // namespace

extern "C" {
int ref_barycentric_tet(double vol, const double M[12], const double p[3], double bcc[4]) {
  o::Vector<4> b;
  const bool ok = pumipic::barycentric_tet(vol, tet(M), v3(p), b);
  for (int i = 0; i < 4; ++i) bcc[i] = b[i];
  return ok;
}
void ref_barycentric_tri(double area, const double M[6], const double p[2], double bcc[3]) {
  o::Matrix<2, 3> m;
  for (int i = 0; i < 3; ++i) m[i] = v2(M + 2 * i);
  o::Vector<3> b;
  pumipic::barycentric_tri(area, m, v2(p), b);
  for (int i = 0; i < 3; ++i) bcc[i] = b[i];
}
int ref_find_barycentric_tet(const double M[12], const double p[3], double bcc[4]) {
  o::Vector<4> b;
  const bool ok = pumipic::find_barycentric_tet(tet(M), v3(p), b);
  for (int i = 0; i < 4; ++i) bcc[i] = b[i];
  return ok;
}
int ref_all_positive(const double* v, int n, double tol) {
  if (n == 3) { o::Vector<3> a; for (int i = 0; i < 3; ++i) a[i] = v[i]; return pumipic::all_positive(a, tol); }
  o::Vector<4> a;
  for (int i = 0; i < 4; ++i) a[i] = v[i];
  return pumipic::all_positive(a, tol);
}
int ref_min_index(const double* v, int n) { return pumipic::min_index(v, n); }
int ref_max_index(const double* v, int n) { return pumipic::max_index(v, n); }
int ref_min3(const double v[3]) { o::Vector<3> a; for (int i = 0; i < 3; ++i) a[i] = v[i]; return pumipic::min3(a); }
int ref_is_face_flipped_3d(int fi, const int fv[3], const int tv[4]) {
  o::Few<o::LO, 3> f; o::Few<o::LO, 4> t;
  for (int i = 0; i < 3; ++i) f[i] = fv[i];
  for (int i = 0; i < 4; ++i) t[i] = tv[i];
  return pumipic::isFaceFlipped(fi, f, t);
}
int ref_is_face_flipped_2d(int ei, const int ev[2], const int tv[3]) {
  o::Few<o::LO, 2> e; o::Few<o::LO, 3> t;
  for (int i = 0; i < 2; ++i) e[i] = ev[i];
  for (int i = 0; i < 3; ++i) t[i] = tv[i];
  return pumipic::isFaceFlipped(ei, e, t);
}
int ref_ray_intersects_triangle(const double face[9], const double orig[3], const double dest[3],
                                double xpoint[3], double tol, int flip, double* dproj,
                                double* closeness, double* param) {
  o::Vector<3> xp = o::zero_vector<3>();
  const bool hit = pumipic::ray_intersects_triangle(tri3(face), v3(orig), v3(dest), xp, tol, flip, *dproj,
                                                    *closeness, *param);
  for (int i = 0; i < 3; ++i) xpoint[i] = xp[i];
  return hit;
}
int ref_line_segment_intersects_triangle(const double face[9], const double orig[3], const double dest[3],
                                         double xpoint[3], double tol, int flip, double* dproj,
                                         double* closeness, double* param) {
  o::Vector<3> xp = o::zero_vector<3>();
  const bool hit = pumipic::line_segment_intersects_triangle(tri3(face), v3(orig), v3(dest), xp, tol, flip,
                                                             *dproj, *closeness, *param);
  for (int i = 0; i < 3; ++i) xpoint[i] = xp[i];
  return hit;
}
int ref_line_edge_2d(const double edge[4], const double orig[2], const double dest[2], double xpoint[2],
                     double tol, int flip) {
  o::Few<o::Vector<2>, 2> e;
  e[0] = v2(edge); e[1] = v2(edge + 2);
  o::Vector<2> xp = o::zero_vector<2>();
  const bool hit = pumipic::line_edge_2d(e, v2(orig), v2(dest), xp, tol, flip);
  xpoint[0] = xp[0]; xpoint[1] = xp[1];
  return hit;
}
int ref_line_triangle_intx_simple(const double abc[9], const double origin[3], const double dest[3],
                                  double xpoint[3], double* dproj, int reverse, double tol) {
  o::Vector<3> xp;
  const bool hit = pumipic::line_triangle_intx_simple(tri3(abc), v3(origin), v3(dest), xp, *dproj, reverse != 0, tol);
  for (int i = 0; i < 3; ++i) xpoint[i] = xp[i];
  return hit;
}
int ref_find_exit_face_bcc_3d(double vol, const double M[12], const double p[3], int* done) {
  o::LO d = 0;
  const int f = pumipic::find_exit_face_bcc_3d(vol, tet(M), v3(p), d);
  *done = d;
  return f;
}
}  // extern "C"

// ---- the reference's search_mesh (adjacency.tpp:642 -> trace_particle_through_mesh :461) on a mesh
// given by its arrays; derived arrays (ask_up, exposed sides, element measures) come from the caller
struct Seg3 {
  const double* p; long stride;
  double operator()(int pid, int i) const { return p[(long)i * stride + pid]; }
};
struct SegI {
  const int* p;
  int operator()(int pid) const { return p[pid]; }
};
template <class T> static o::Write<T> to_write(const T* a, long n) {
  o::Write<T> w((int)n, T());
  for (long i = 0; i < n; ++i) w[(int)i] = a[i];
  return w;
}
typedef pumipic::MemberTypes<int> RefParticle;   // the searches never touch member data

extern "C" int ref_search_mesh(int dim, int nverts, const double* coords, int nelems, const int* elem2verts,
                    int nsides, const int* elem2sides, const int* side2verts, const int* side2elem_off,
                    const int* side2elem, const signed char* exposed, const double* measure, int cap,
                    const int* slot_elem, const unsigned char* mask, const double* x, const double* xtgt,
                    long stride, int* elem_ids, int elem_ids_empty, int require_intersection,
                    int* inter_faces, double* inter_points, int inter_given, int looplimit) {
  o::Mesh mesh;
  mesh.dim_ = dim;
  mesh.coords_ = o::Reals(to_write(coords, (long)nverts * dim));
  mesh.elem_verts = o::LOs(to_write(elem2verts, (long)nelems * (dim + 1)));
  mesh.down = o::LOs(to_write(elem2sides, (long)nelems * (dim + 1)));
  mesh.side_verts = o::LOs(to_write(side2verts, (long)nsides * dim));
  mesh.up_off = o::LOs(to_write(side2elem_off, (long)nsides + 1));
  mesh.up_vals = o::LOs(to_write(side2elem, (long)side2elem_off[nsides]));
  mesh.exposed = o::Bytes(to_write(exposed, (long)nsides));
  mesh.measure = o::Reals(to_write(measure, (long)nelems));
  pumipic::ParticleStructure<RefParticle> ptcls;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask;
  std::vector<int> pid((size_t)cap);
  for (int i = 0; i < cap; ++i) pid[(size_t)i] = i;
  Seg3 xo{x, stride}, xt{xtgt, stride};
  SegI pids{pid.data()};
  o::Write<o::LO> ids, faces;
  o::Write<o::Real> pts;
  if (!elem_ids_empty) ids = to_write(elem_ids, cap);
  if (inter_given) { faces = to_write(inter_faces, cap); pts = to_write(inter_points, (long)dim * cap); }
  const bool found = pumipic::search_mesh(mesh, &ptcls, xo, xt, pids, ids, require_intersection != 0, faces, pts,
                                          looplimit, 0);
  for (int i = 0; i < cap; ++i) elem_ids[i] = ids[i];
  if (require_intersection) {
    for (int i = 0; i < cap; ++i) inter_faces[i] = faces[i];
    for (long i = 0; i < (long)dim * cap; ++i) inter_points[i] = pts[(int)i];
  }
  return found;
}

static void fill_mesh(o::Mesh& mesh, int dim, int nverts, const double* coords, int nelems, const int* elem2verts,
                      int nsides, const int* elem2sides, const int* side2verts, const int* side2elem_off,
                      const int* side2elem, const signed char* exposed, const double* measure) {
  mesh.dim_ = dim;
  mesh.coords_ = o::Reals(to_write(coords, (long)nverts * dim));
  mesh.elem_verts = o::LOs(to_write(elem2verts, (long)nelems * (dim + 1)));
  mesh.down = o::LOs(to_write(elem2sides, (long)nelems * (dim + 1)));
  mesh.side_verts = o::LOs(to_write(side2verts, (long)nsides * dim));
  mesh.up_off = o::LOs(to_write(side2elem_off, (long)nsides + 1));
  mesh.up_vals = o::LOs(to_write(side2elem, (long)side2elem_off[nsides]));
  mesh.exposed = o::Bytes(to_write(exposed, (long)nsides));
  mesh.measure = o::Reals(to_write(measure, (long)nelems));
}

// the reference's search_mesh_2d (adjacency.hpp:1013-1158): elem_ids in/out (-1 = use the row element)
extern "C" int ref_search_mesh_2d(int nverts, const double* coords, int nelems, const int* elem2verts, int nsides,
                                  const int* elem2sides, const int* side2verts, const int* side2elem_off,
                                  const int* side2elem, const signed char* exposed, const double* measure,
                                  int cap, const int* slot_elem, const unsigned char* mask, const double* x,
                                  const double* xtgt, long stride, int* elem_ids, int looplimit) {
  o::Mesh mesh;
  fill_mesh(mesh, 2, nverts, coords, nelems, elem2verts, nsides, elem2sides, side2verts, side2elem_off, side2elem,
            exposed, measure);
  pumipic::ParticleStructure<RefParticle> ptcls;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask;
  std::vector<int> pid((size_t)cap);
  for (int i = 0; i < cap; ++i) pid[(size_t)i] = i;
  Seg3 xo{x, stride}, xt{xtgt, stride};
  SegI pids{pid.data()};
  o::Write<o::LO> ids = to_write(elem_ids, cap);
  const bool found = pumipic::search_mesh_2d(mesh, &ptcls, xo, xt, pids, ids, looplimit, false);
  for (int i = 0; i < cap; ++i) elem_ids[i] = ids[i];
  return found;
}

// the reference's legacy 3D search_mesh (adjacency.hpp:559-768; variant 0) and search_mesh_3d
// (:316-555; variant 1).  Returns found (0/1), or -2 when an OMEGA_H_CHECK of the reference fired
// (it would have aborted the process: a particle's origin is not in its start element).
extern "C" int ref_search_mesh_3d_variants(int variant, int nverts, const double* coords, int nelems,
                                           const int* elem2verts, int nsides, const int* elem2sides,
                                           const int* side2verts, const int* side2elem_off, const int* side2elem,
                                           const signed char* exposed, const double* measure, const int* dual_off,
                                           const int* dual, int cap, const int* slot_elem, const unsigned char* mask,
                                           const double* x, const double* xtgt, long stride, int* elem_ids,
                                           int elem_ids_empty, double* xpoints, int* xface, int looplimit) {
  o::Mesh mesh;
  fill_mesh(mesh, 3, nverts, coords, nelems, elem2verts, nsides, elem2sides, side2verts, side2elem_off, side2elem,
            exposed, measure);
  mesh.dual_off = o::LOs(to_write(dual_off, (long)nelems + 1));
  mesh.dual_vals = o::LOs(to_write(dual, (long)dual_off[nelems]));
  pumipic::ParticleStructure<RefParticle> ptcls;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask;
  std::vector<int> pid((size_t)cap);
  for (int i = 0; i < cap; ++i) pid[(size_t)i] = i;
  Seg3 xo{x, stride}, xt{xtgt, stride};
  SegI pids{pid.data()};
  o::Write<o::LO> ids;
  if (!elem_ids_empty) ids = to_write(elem_ids, cap);
  o::Write<o::Real> xp = to_write(xpoints, 3L * cap);
  o::Write<o::LO> xf = to_write(xface, cap);
  int found;
  try {
    if (variant == 0) found = pumipic::search_mesh(mesh, &ptcls, xo, xt, pids, ids, xp, xf, looplimit, 0);
    else found = pumipic::search_mesh_3d(mesh, &ptcls, xo, xt, pids, ids, xp, xf, looplimit, 0);
  } catch (const RefCheckFailed&) {
    return -2;
  }
  for (int i = 0; i < cap; ++i) { elem_ids[i] = ids[i]; xface[i] = xf[i]; }
  for (long i = 0; i < 3L * cap; ++i) xpoints[i] = xp[(int)i];
  return found;
}

// ---- gather helpers (adjacency.hpp:772-809, utils.hpp:245-456); -2 = an OMEGA_H_CHECK fired
extern "C" int ref_interpolate_tet_vtx(const int* elem2verts, int nelems, const double* field, long nfield, int elem,
                                       const double bcc[4], int dof, int comp, double* out) {
  o::Vector<4> b;
  for (int i = 0; i < 4; ++i) b[i] = bcc[i];
  try {
    *out = pumipic::interpolateTetVtx(o::LOs(to_write(elem2verts, 4L * nelems)), o::Reals(to_write(field, nfield)),
                                      elem, b, dof, comp);
  } catch (const RefCheckFailed&) { return -2; }
  return 0;
}
extern "C" int ref_find_bcc_in_tet(const double* coords, int nverts, const int* elem2verts, int nelems,
                                   const double xyz[3], int elem, double bcc[4]) {
  o::Vector<4> b;
  try {
    pumipic::findBCCoordsInTet(o::Reals(to_write(coords, 3L * nverts)), o::LOs(to_write(elem2verts, 4L * nelems)),
                               v3(xyz), elem, b);
  } catch (const RefCheckFailed&) { return -2; }
  for (int i = 0; i < 4; ++i) bcc[i] = b[i];
  return 0;
}
extern "C" double ref_interpolate2d_field(const double* data, long ndata, double gridx0, double gridz0, double dx,
                                          double dz, int nx, int nz, const double pos[3], int cyl, int nComp, int comp) {
  return pumipic::interpolate2d_field(o::Reals(to_write(data, ndata)), gridx0, gridz0, dx, dz, nx, nz, v3(pos), cyl != 0,
                                      nComp, comp);
}
extern "C" void ref_interp2d_vector(const double* data3, long ndata, double gridx0, double gridz0, double dx, double dz,
                                    int nx, int nz, const double pos[3], double field[3], int cyl) {
  o::Vector<3> f = o::zero_vector<3>();
  pumipic::interp2dVector(o::Reals(to_write(data3, ndata)), gridx0, gridz0, dx, dz, nx, nz, v3(pos), f, cyl != 0);
  for (int i = 0; i < 3; ++i) field[i] = f[i];
}
extern "C" double ref_interpolate3d_field(double x, double y, double z, int nx, int ny, int nz, const double* gridx,
                                          const double* gridy, const double* gridz, const double* data) {
  return pumipic::interpolate3d_field(x, y, z, nx, ny, nz, o::Reals(to_write(gridx, nx)), o::Reals(to_write(gridy, ny)),
                                      o::Reals(to_write(gridz, nz)), o::Reals(to_write(data, (long)nx * ny * nz)));
}

// threads of the OpenMP loops of the stand-ins (1 = the serial order the oracle uses)
extern "C" void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
extern "C" int ref_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
