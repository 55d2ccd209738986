// omega_h_shim.hpp -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).
//
// Just enough of Omega_h's and Kokkos' small-vector vocabulary for the reference's own geometric
// primitives (src/pumipic_adjacency.tpp, src/pumipic_adjacency.hpp, src/pumipic_utils.hpp) to
// compile UNMODIFIED outside their build system: oracle/build_ref_primitives.py extracts those
// functions from /root/reference at build time (into the git-ignored oracle/_ref/) and compiles
// them against this header.  Nothing here comes from the reference tree; it restates the
// published definitions of the third-party pieces the primitives call (SCOREC/omega_h, CI pin
// scorec-v10.8.4: Omega_h_few.hpp, Omega_h_vector.hpp, Omega_h_matrix.hpp, Omega_h_simplex.hpp,
// Omega_h_scalar.hpp), with the same evaluation order:
//   inner_product: a[0]*b[0], then += a[i]*b[i] in index order;  norm = sqrt(inner_product(a,a))
//   cross(3D): (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0);    cross(2D): a0*b1 - a1*b0
//   normalize(v) = v / norm(v);  perp(v) = (-v1, v0);  triangle_area_from_basis(b) = cross(b0,b1)/2
//   are_close(a,b,tol,floor): |a|,|b| <= floor, else |b-a| / max(|a|,|b|) <= tol
//   Matrix<m,n> = Few<Vector<m>, n> (n column vectors); simplex_down_template as in Omega_h_simplex.hpp
#pragma once
#include <cassert>
#include <cmath>

#define OMEGA_H_DEVICE static inline
#define OMEGA_H_INLINE static inline
// OMEGA_H_CHECK aborts the reference; here it throws so that a wrapper can report "would have aborted"
struct RefCheckFailed {};
#define OMEGA_H_CHECK(cond) do { if (!(cond)) throw RefCheckFailed(); } while (0)
#define printInfo(...) ((void)0)

namespace Kokkos {
static inline double fabs(double a) { return std::fabs(a); }
static inline double abs(double a) { return std::fabs(a); }
static inline double min(double a, double b) { return (b < a) ? b : a; }   // std::min
static inline double max(double a, double b) { return (a < b) ? b : a; }   // std::max
static inline double floor(double a) { return std::floor(a); }
static inline double cos(double a) { return std::cos(a); }
static inline double sin(double a) { return std::sin(a); }
static inline double sqrt(double a) { return std::sqrt(a); }
static inline double pow(double a, int b) { return std::pow(a, b); }
static inline double pow(double a, double b) { return std::pow(a, b); }
static inline double atan2(double a, double b) { return std::atan2(a, b); }
}  // namespace Kokkos

namespace Omega_h {
typedef double Real;
typedef int LO;
enum { VERT = 0, EDGE = 1, FACE = 2, REGION = 3 };

template <typename T, int n>
class Few {
  T array_[n];

 public:
  Few() {}
  int size() const { return n; }
  T& operator[](int i) { return array_[i]; }
  const T& operator[](int i) const { return array_[i]; }
};

template <int n>
class Vector : public Few<Real, n> {
 public:
  Vector() {}
};

template <int m, int n>
class Matrix : public Few<Vector<m>, n> {
 public:
  Matrix() {}
};

template <int n> Vector<n> operator+(Vector<n> a, Vector<n> b) { Vector<n> c; for (int i = 0; i < n; ++i) c[i] = a[i] + b[i]; return c; }
template <int n> Vector<n> operator-(Vector<n> a, Vector<n> b) { Vector<n> c; for (int i = 0; i < n; ++i) c[i] = a[i] - b[i]; return c; }
template <int n> Vector<n> operator*(Vector<n> a, Real b) { Vector<n> c; for (int i = 0; i < n; ++i) c[i] = a[i] * b; return c; }
template <int n> Vector<n> operator*(Real a, Vector<n> b) { return b * a; }
template <int n> Vector<n> operator/(Vector<n> a, Real b) { Vector<n> c; for (int i = 0; i < n; ++i) c[i] = a[i] / b; return c; }
template <int n> Vector<n> zero_vector() { Vector<n> v; for (int i = 0; i < n; ++i) v[i] = 0.0; return v; }

template <int n> Real inner_product(Vector<n> a, Vector<n> b) {
  Real out = a[0] * b[0];
  for (int i = 1; i < n; ++i) out += a[i] * b[i];
  return out;
}
template <int n> Real norm_squared(Vector<n> v) { return inner_product(v, v); }
template <int n> Real norm(Vector<n> v) { return std::sqrt(norm_squared(v)); }
template <int n> Vector<n> normalize(Vector<n> v) { return v / norm(v); }

static inline Vector<3> cross(Vector<3> a, Vector<3> b) {
  Vector<3> c;
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  return c;
}
static inline Real cross(Vector<2> a, Vector<2> b) { return a[0] * b[1] - a[1] * b[0]; }
static inline Vector<2> perp(Vector<2> v) { Vector<2> r; r[0] = -v[1]; r[1] = v[0]; return r; }
static inline Real triangle_area_from_basis(Few<Vector<2>, 2> b) { return cross(b[0], b[1]) / 2.0; }
// simplex_basis<3,3>: edge vectors p[i+1] - p[0];  tet_volume_from_basis(b) = (cross(b0,b1) . b2) / 6
template <int sdim, int edim> Few<Vector<sdim>, edim> simplex_basis(Few<Vector<sdim>, edim + 1> p) {
  Few<Vector<sdim>, edim> b;
  for (int i = 0; i < edim; ++i) b[i] = p[i + 1] - p[0];
  return b;
}
static inline Real tet_volume_from_basis(Few<Vector<3>, 3> b) { return inner_product(cross(b[0], b[1]), b[2]) / 6.0; }

static inline Real rel_diff_with_floor(Real a, Real b, Real floor) {
  Real am = std::fabs(a), bm = std::fabs(b);
  if (am <= floor && bm <= floor) return 0.0;
  return std::fabs(b - a) / ((am < bm) ? bm : am);
}
static inline bool are_close(Real a, Real b, Real tol = 1e-10, Real floor = 1e-10) {
  return rel_diff_with_floor(a, b, floor) <= tol;
}

static inline int simplex_down_template(int elem_dim, int bdry_dim, int which_bdry, int which_vert) {
  static const int tri_edges[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int tet_edges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  static const int tet_faces[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};
  if (bdry_dim == 0) return which_bdry;
  if (elem_dim == 1) return which_vert;
  if (elem_dim == 2) return tri_edges[which_bdry][which_vert];
  if (bdry_dim == 1) return tet_edges[which_bdry][which_vert];
  return tet_faces[which_bdry][which_vert];
}
// vertex opposite to a side: triangle edge i -> (i + 2) % 3; tet face 0,1,2,3 -> 3,2,0,1
static inline int simplex_opposite_template(int elem_dim, int bdry_dim, int which_bdry) {
  static const int tet_opp[4] = {3, 2, 0, 1};
  if (elem_dim == 3 && bdry_dim == 2) return tet_opp[which_bdry];
  if (elem_dim == 2 && bdry_dim == 1) return (which_bdry + 2) % 3;
  return -1;
}
}  // namespace Omega_h
