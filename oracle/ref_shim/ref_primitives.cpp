// ref_primitives.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" entry points, with the same
// signatures as the oracle's primitives (oracle/pumipic_oracle.h:60-83), around the reference's own
// functions: their text is pulled in from oracle/_ref/ref_primitives.inc, which
// oracle/build_ref_primitives.py extracts from /root/reference at build time.
#include "omega_h_shim.hpp"

namespace o = Omega_h;
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
}
