/*
 * pumipic_oracle.c -- CPU restatement of PUMI-PIC's particle hot path (see pumipic_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product library.
 *
 * Structure mirrors the reference: one loop nest per reference kernel ("kernel-per-phase"),
 * each loop an `omp parallel for` over all slots of the particle structure, with a
 * host-side reduction after every walk iteration exactly where the reference does
 * `o::get_min(ptcl_done)`.  Compile with -ffp-contract=off so no FMA contraction occurs
 * (the reference's CPU builds on baseline x86-64 have none either).
 */
#include "pumipic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPSILON 1e-10 /* src/pumipic_constants.hpp:6 */

/* ------------------------------------------------------------------------------------------
 * Omega_h small-vector arithmetic (third-party, not vendored: SCOREC/omega_h scorec-v10.8.4,
 * Omega_h_vector.hpp / Omega_h_shape.hpp).  Restated from the published formulas; call sites
 * in the reference: adjacency.tpp:36,56,161-176,211-215; adjacency.hpp:110,117,169-180,242-256.
 * ---------------------------------------------------------------------------------------- */
static inline void v3_sub(const double a[3], const double b[3], double c[3]) {
  c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2];
}
static inline void v3_cross(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double v3_dot(const double a[3], const double b[3]) {
  double c = a[0] * b[0];
  c = c + a[1] * b[1];
  c = c + a[2] * b[2];
  return c;
}
static inline double v3_norm(const double a[3]) { return sqrt(v3_dot(a, a)); }
static inline double v2_dot(const double a[2], const double b[2]) {
  double c = a[0] * b[0];
  c = c + a[1] * b[1];
  return c;
}
static inline double v2_cross(const double a[2], const double b[2]) {
  return a[0] * b[1] - a[1] * b[0];
}
/* Omega_h::are_close(a, b, tol, floor) */
static inline int are_close(double a, double b, double tol, double floor_) {
  double am = fabs(a), bm = fabs(b);
  if (am <= floor_ && bm <= floor_) return 1;
  double mx = am > bm ? am : bm;
  return fabs(b - a) / mx <= tol;
}

/* Omega_h simplex_down_template(3,2,f,i): tet faces (0,2,1),(0,1,3),(1,2,3),(2,0,3) */
static const int TET_FACE[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};
/* simplex_down_template(2,1,e,i): triangle edges (0,1),(1,2),(2,0) */
static const int TRI_EDGE[3][2] = {{0, 1}, {1, 2}, {2, 0}};

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Mesh derivations (Omega_h: ask_up, mark_exposed_sides, ask_dual, measure_elements_real).
 * Used by adjacency.tpp:489,497-501 and adjacency.hpp:568-574,1030-1036.
 * ---------------------------------------------------------------------------------------- */
orc_mesh* orc_mesh_create(int dim, int nverts, const double* coords, int nelems,
                          const int* elem2verts, int nsides, const int* elem2sides,
                          const int* side2verts) {
  orc_mesh* m = (orc_mesh*)calloc(1, sizeof(orc_mesh));
  const int nv = dim + 1;
  m->dim = dim; m->nverts = nverts; m->nelems = nelems; m->nsides = nsides;
  m->coords = coords; m->elem2verts = elem2verts; m->elem2sides = elem2sides;
  m->side2verts = side2verts;
  /* ask_up(dim-1, dim): adjacent elements of every side in ascending id order */
  m->side2elem_off = (int*)calloc((size_t)nsides + 1, sizeof(int));
  for (long i = 0; i < (long)nelems * nv; ++i) m->side2elem_off[elem2sides[i] + 1]++;
  for (int s = 0; s < nsides; ++s) m->side2elem_off[s + 1] += m->side2elem_off[s];
  m->side2elem = (int*)malloc(sizeof(int) * (size_t)m->side2elem_off[nsides]);
  int* fill = (int*)calloc((size_t)nsides, sizeof(int));
  for (int e = 0; e < nelems; ++e)
    for (int k = 0; k < nv; ++k) {
      int s = elem2sides[(long)e * nv + k];
      m->side2elem[m->side2elem_off[s] + fill[s]++] = e;
    }
  free(fill);
  /* mark_exposed_sides: exactly one upward-adjacent element */
  m->exposed = (signed char*)malloc((size_t)nsides);
  for (int s = 0; s < nsides; ++s)
    m->exposed[s] = (m->side2elem_off[s + 1] - m->side2elem_off[s]) == 1;
  /* ask_dual: neighbour across each non-exposed side, in local side order */
  m->dual_off = (int*)calloc((size_t)nelems + 1, sizeof(int));
  for (int e = 0; e < nelems; ++e) {
    int c = 0;
    for (int k = 0; k < nv; ++k) c += !m->exposed[elem2sides[(long)e * nv + k]];
    m->dual_off[e + 1] = m->dual_off[e] + c;
  }
  m->dual = (int*)malloc(sizeof(int) * (size_t)(m->dual_off[nelems] + 1));
  for (int e = 0; e < nelems; ++e) {
    int p = m->dual_off[e];
    for (int k = 0; k < nv; ++k) {
      int s = elem2sides[(long)e * nv + k];
      if (m->exposed[s]) continue;
      int a = m->side2elem[m->side2elem_off[s]];
      int b = m->side2elem[m->side2elem_off[s] + 1];
      m->dual[p++] = (a == e) ? b : a;
    }
  }
  /* measure_elements_real: triangle_area_from_basis / tet_volume_from_basis of simplex_basis */
  m->vol = (double*)malloc(sizeof(double) * (size_t)nelems);
  for (int e = 0; e < nelems; ++e) {
    const int* v = elem2verts + (long)e * nv;
    if (dim == 2) {
      const double* p0 = coords + 2 * (long)v[0];
      const double* p1 = coords + 2 * (long)v[1];
      const double* p2 = coords + 2 * (long)v[2];
      double b0[2] = {p1[0] - p0[0], p1[1] - p0[1]};
      double b1[2] = {p2[0] - p0[0], p2[1] - p0[1]};
      m->vol[e] = v2_cross(b0, b1) / 2.0;
    } else {
      const double* p0 = coords + 3 * (long)v[0];
      double b0[3], b1[3], b2[3], c[3];
      v3_sub(coords + 3 * (long)v[1], p0, b0);
      v3_sub(coords + 3 * (long)v[2], p0, b1);
      v3_sub(coords + 3 * (long)v[3], p0, b2);
      v3_cross(b0, b1, c);
      m->vol[e] = v3_dot(c, b2) / 6.0;
    }
  }
  /* ask_up(0, dim) */
  m->vert2elem_off = (int*)calloc((size_t)nverts + 1, sizeof(int));
  for (long i = 0; i < (long)nelems * nv; ++i) m->vert2elem_off[elem2verts[i] + 1]++;
  for (int v = 0; v < nverts; ++v) m->vert2elem_off[v + 1] += m->vert2elem_off[v];
  m->vert2elem = (int*)malloc(sizeof(int) * (size_t)m->vert2elem_off[nverts]);
  fill = (int*)calloc((size_t)nverts, sizeof(int));
  for (int e = 0; e < nelems; ++e)
    for (int k = 0; k < nv; ++k) {
      int v = elem2verts[(long)e * nv + k];
      m->vert2elem[m->vert2elem_off[v] + fill[v]++] = e;
    }
  free(fill);
  return m;
}

void orc_mesh_destroy(orc_mesh* m) {
  if (!m) return;
  free(m->side2elem_off); free(m->side2elem); free(m->dual_off); free(m->dual);
  free(m->exposed); free(m->vol); free(m->vert2elem_off); free(m->vert2elem);
  free(m);
}

/* adjacency.tpp:419-428 compute_tolerance_from_area */
double orc_compute_tolerance(const orc_mesh* m) {
  double min_area = INFINITY; /* Kokkos::Min identity */
  for (int e = 0; e < m->nelems; ++e)
    if (m->vol[e] < min_area) min_area = m->vol[e];
  double t = 1e-15 / min_area;
  return t > 1e-8 ? t : 1e-8;
}

/* ------------------------------------------------------------------------------------------
 * Geometry primitives
 * ---------------------------------------------------------------------------------------- */

/* pumipic_utils.hpp:78-86 all_positive */
int orc_all_positive(const double* v, int n, double tol) {
  int pos = 1;
  for (int i = 0; i < n; ++i) {
    int gtez = are_close(v[i], 0.0, tol, tol) || v[i] > 0;
    pos = pos && gtez;
  }
  return pos;
}
/* pumipic_utils.hpp:88-92 min3 */
int orc_min3(const double a[3]) {
  int idx = (a[0] < a[1]) ? 0 : 1;
  idx = (a[idx] < a[2]) ? idx : 2;
  return idx;
}
/* pumipic_utils.hpp:125-136 min_index */
int orc_min_index(const double* a, int n) {
  int ind = 0;
  double mn = a[0];
  for (int i = 0; i < n - 1; ++i)
    if (mn > a[i + 1]) { mn = a[i + 1]; ind = i + 1; }
  return ind;
}
/* pumipic_utils.hpp:138-149 max_index (beg = 0) */
int orc_max_index(const double* a, int n) {
  int ind = 0;
  double mx = a[0];
  for (int i = 0; i < n - 1; ++i)
    if (mx < a[i + 1]) { mx = a[i + 1]; ind = i + 1; }
  return ind;
}

/* adjacency.tpp:41-69 barycentric_tet (+ utils.hpp:565-572 get_face_from_face_index_of_tet).
 * M = 4 vertices x 3 coords, row-major. */
int orc_barycentric_tet(double vol, const double M[12], const double p[3], double bcc[4]) {
  double vals[4];
  for (int i = 0; i < 4; ++i) bcc[i] = -1;
  for (int f = 0; f < 4; ++f) {
    const double* a = M + 3 * TET_FACE[f][0];
    const double* b = M + 3 * TET_FACE[f][1];
    const double* c = M + 3 * TET_FACE[f][2];
    double vab[3], vac[3], vap[3], cr[3];
    v3_sub(b, a, vab);
    v3_sub(c, a, vac);
    v3_sub(p, a, vap);
    v3_cross(vac, vab, cr);
    vals[f] = v3_dot(vap, cr);
  }
  double inv_vol = 0.0;
  if (vol > 0) inv_vol = 1.0 / vol;
  else return 0;
  for (int i = 0; i < 4; ++i) bcc[i] = inv_vol * vals[i];
  return 1;
}

/* adjacency.tpp:23-39 barycentric_tri (same arithmetic as adjacency.hpp:75-94).
 * M = 3 vertices x 2 coords. */
void orc_barycentric_tri(double area, const double M[6], const double p[2], double bcc[3]) {
  for (int i = 0; i < 3; ++i) {
    const double* k = M + 2 * TRI_EDGE[i][0];
    const double* l = M + 2 * TRI_EDGE[i][1];
    double t0[2] = {l[0] - k[0], l[1] - k[1]};
    double t1[2] = {p[0] - k[0], p[1] - k[1]};
    double a = v2_cross(t0, t1) / 2.0; /* triangle_area_from_basis */
    bcc[i] = a / area;
  }
}

/* adjacency.hpp:97-133 find_barycentric_tet (legacy: own vol6, sums to 1) */
int orc_find_barycentric_tet(const double M[12], const double p[3], double bcc[4]) {
  double vals[4];
  for (int i = 0; i < 4; ++i) bcc[i] = -1;
  for (int f = 0; f < 4; ++f) {
    const double* a = M + 3 * TET_FACE[f][0];
    const double* b = M + 3 * TET_FACE[f][1];
    const double* c = M + 3 * TET_FACE[f][2];
    double vab[3], vac[3], vap[3], cr[3];
    v3_sub(b, a, vab);
    v3_sub(c, a, vac);
    v3_sub(p, a, vap);
    v3_cross(vac, vab, cr);
    vals[f] = v3_dot(vap, cr);
  }
  /* volume from the bottom face: abc = M0, M2, M1; cross(abc2-abc0, abc1-abc0) */
  const double* a0 = M + 3 * TET_FACE[0][0];
  const double* a1 = M + 3 * TET_FACE[0][1];
  const double* a2 = M + 3 * TET_FACE[0][2];
  double e2[3], e1[3], cr[3], d3[3];
  v3_sub(a2, a0, e2);
  v3_sub(a1, a0, e1);
  v3_cross(e2, e1, cr);
  v3_sub(M + 9, M, d3);
  double vol6 = v3_dot(d3, cr);
  if (!(vol6 > 1.0e-20)) return 0;
  double inv = 1.0 / vol6;
  for (int i = 0; i < 4; ++i) bcc[i] = inv * vals[i];
  return 1;
}

/* pumipic_utils.hpp:489-493 getFaceMap */
static const int FACE_MAP[8] = {2, 1, 1, 3, 2, 3, 0, 3};

/* pumipic_utils.hpp:501-507 isFaceFlipped (3D) */
int orc_is_face_flipped_3d(int fi, const int fv[3], const int tv[4]) {
  int m1 = FACE_MAP[fi * 2], m2 = FACE_MAP[fi * 2 + 1];
  int idx = (fv[0] == tv[m1]) ? 1 : (fv[1] == tv[m1]) ? 2 : 0;
  return tv[m2] != fv[idx];
}
/* pumipic_utils.hpp:495-499 isFaceFlipped (2D) */
int orc_is_face_flipped_2d(const int ev[2], const int tv[3]) {
  int idx = (ev[0] == tv[0]) ? 1 : (ev[0] == tv[1]) ? 2 : 0;
  return ev[1] != tv[idx];
}

/* Kokkos::min / Kokkos::max follow std::min / std::max: (b<a)?b:a and (a<b)?b:a */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }

/* adjacency.tpp:152-178 ray_intersects_triangle.  face = 3 vertices x 3 coords */
int orc_ray_intersects_triangle(const double face[9], const double orig[3], const double dest[3],
                                double xpoint[3], double tol, int flip, double* dproj_out,
                                double* closeness_out, double* param_out) {
  const int vtx1 = 2 - flip, vtx2 = flip + 1;
  double edge1[3], edge2[3], disp[3], dir[3], fnorm[3], pvec[3], tvec[3], qvec[3];
  v3_sub(face + 3 * vtx1, face, edge1);
  v3_sub(face + 3 * vtx2, face, edge2);
  v3_sub(dest, orig, disp);
  const double seg_length = v3_norm(disp);
  for (int i = 0; i < 3; ++i) dir[i] = disp[i] / seg_length;
  v3_cross(edge2, edge1, fnorm);
  v3_cross(dir, edge2, pvec);
  const double dproj = v3_dot(dir, fnorm);
  const double invdet = 1.0 / dproj;
  v3_sub(orig, face, tvec);
  const double u = invdet * v3_dot(tvec, pvec);
  v3_cross(tvec, edge1, qvec);
  const double v = invdet * v3_dot(dir, qvec);
  const double t = invdet * v3_dot(edge2, qvec);
  *param_out = t / seg_length;
  for (int i = 0; i < 3; ++i) xpoint[i] = orig[i] + dir[i] * t;
  *closeness_out = dmax(dmax(dmin(fabs(u), fabs(1 - u)), dmin(fabs(v), fabs(1 - v))),
                        dmin(fabs(u + v), fabs(1 - u - v)));
  *dproj_out = dproj;
  return (dproj >= tol) && (t >= -tol) && (u >= -tol) && (v >= -tol) &&
         (u + v <= 1.0 + 2 * tol);
}

/* adjacency.tpp:192-201 */
int orc_line_segment_intersects_triangle(const double face[9], const double orig[3],
                                         const double dest[3], double xpoint[3], double tol,
                                         int flip, double* dproj, double* closeness,
                                         double* param) {
  int hit = orc_ray_intersects_triangle(face, orig, dest, xpoint, tol, flip, dproj, closeness,
                                        param);
  return hit && *param <= 1 + tol;
}

/* adjacency.tpp:204-218 line_edge_2d.  edge = 2 vertices x 2 coords */
int orc_line_edge_2d(const double edge[4], const double orig[2], const double dest[2],
                     double xpoint[2], double tol, int flip) {
  const int vtx1 = flip, vtx2 = !flip;
  const double path[2] = {dest[0] - orig[0], dest[1] - orig[1]};
  const double ed[2] = {edge[2 * vtx2] - edge[2 * vtx1], edge[2 * vtx2 + 1] - edge[2 * vtx1 + 1]};
  const double nrm[2] = {-ed[1], ed[0]};      /* perp(edge) */
  const double nrmp[2] = {-path[1], path[0]}; /* perp(path) */
  const double det = -v2_dot(nrm, path);
  const double rel[2] = {orig[0] - edge[2 * vtx1], orig[1] - edge[2 * vtx1 + 1]};
  const double s = v2_dot(nrmp, rel);
  const double t = v2_dot(nrm, rel);
  const double r = t / det;
  xpoint[0] = orig[0] + r * path[0];
  xpoint[1] = orig[1] + r * path[1];
  return det >= tol && s >= -tol && s <= det + tol && t >= -tol && t <= det + tol;
}

/* adjacency.hpp:163-183 find_barycentric_tri_simple */
static int find_barycentric_tri_simple(const double abc[9], const double xp[3], double bc[3]) {
  const double *a = abc, *b = abc + 3, *c = abc + 6;
  double ba[3], ca[3], cb[3], xa[3], xb[3], cr[3], nrm[3], t[3];
  v3_sub(b, a, ba);
  v3_sub(c, a, ca);
  v3_cross(ba, ca, cr);
  for (int i = 0; i < 3; ++i) cr[i] = cr[i] * (1 / 2.0);
  double len = v3_norm(cr);
  for (int i = 0; i < 3; ++i) nrm[i] = cr[i] / len;
  double area = v3_dot(nrm, cr);
  if (fabs(area) < 1e-20) return 0;
  double fac = 1 / (area * 2.0);
  v3_sub(xp, a, xa);
  v3_cross(ba, xa, t);
  bc[0] = fac * v3_dot(nrm, t);
  v3_sub(c, b, cb);
  v3_sub(xp, b, xb);
  v3_cross(cb, xb, t);
  bc[1] = fac * v3_dot(nrm, t);
  v3_cross(xa, ca, t);
  bc[2] = fac * v3_dot(nrm, t);
  return 1;
}

/* adjacency.hpp:230-273 line_triangle_intx_simple.  *dproj is written only when both plane
 * projections pass, exactly like the reference's by-reference argument. */
int orc_line_triangle_intx_simple(const double abc[9], const double origin[3],
                                  const double dest[3], double xpoint[3], double* dproj,
                                  int reverse, double tol) {
  for (int i = 0; i < 3; ++i) xpoint[i] = 0;
  int found = 0;
  double line[3], edge0[3], edge1[3], normv[3], unit[3], ao[3], p2d[3];
  v3_sub(dest, origin, line);
  v3_sub(abc + 3, abc, edge0);
  v3_sub(abc + 6, abc, edge1);
  v3_cross(edge0, edge1, normv);
  if (reverse)
    for (int i = 0; i < 3; ++i) normv[i] = -1 * normv[i];
  double len = v3_norm(normv);
  for (int i = 0; i < 3; ++i) unit[i] = normv[i] / len;
  v3_sub(abc, origin, ao);
  double dist2plane = v3_dot(ao, unit);
  v3_sub(dest, abc, p2d);
  double proj_end = v3_dot(unit, p2d);
  if (dist2plane >= -tol && proj_end >= -tol) {
    *dproj = v3_dot(line, unit);
    double par_t = (*dproj > 0) ? dist2plane / *dproj : 0;
    for (int i = 0; i < 3; ++i) xpoint[i] = origin[i] + par_t * line[i];
    if (*dproj > 0) {
      double bcc[3];
      int res = find_barycentric_tri_simple(abc, xpoint, bcc);
      if (res && bcc[0] >= 0 && bcc[0] <= 1 && bcc[1] >= 0 && bcc[1] <= 1 && bcc[2] >= 0 &&
          bcc[2] <= 1)
        found = 1;
    }
  }
  return found;
}

/* ------------------------------------------------------------------------------------------
 * helpers for gathering element data
 * ---------------------------------------------------------------------------------------- */
static inline void gather_tet(const orc_mesh* m, int e, int v[4], double M[12]) {
  for (int k = 0; k < 4; ++k) {
    v[k] = m->elem2verts[4 * (long)e + k];
    for (int i = 0; i < 3; ++i) M[3 * k + i] = m->coords[3 * (long)v[k] + i];
  }
}
static inline void gather_tri(const orc_mesh* m, int e, int v[3], double M[6]) {
  for (int k = 0; k < 3; ++k) {
    v[k] = m->elem2verts[3 * (long)e + k];
    for (int i = 0; i < 2; ++i) M[2 * k + i] = m->coords[2 * (long)v[k] + i];
  }
}
static inline void load3(const double* a, long stride, int s, double p[3]) {
  p[0] = a[s]; p[1] = a[stride + s]; p[2] = a[2 * stride + s];
}

static int min_done(const int* done, int cap) {
  int mn = 1; /* o::get_min over the whole capacity */
#pragma omp parallel for reduction(min : mn)
  for (int s = 0; s < cap; ++s)
    if (done[s] < mn) mn = done[s];
  return cap > 0 ? mn : 1;
}

/* ------------------------------------------------------------------------------------------
 * New search API: adjacency.tpp:642 search_mesh -> :461 trace_particle_through_mesh with the
 * RemoveParticleOnGeometricModelExit handler (:618-640).
 * ---------------------------------------------------------------------------------------- */

/* adjacency.tpp:73-145 check_initial_parents */
static int check_initial_parents(const orc_mesh* m, int cap, const unsigned char* mask,
                                 const double* x_orig, long stride, int* elem_ids, int* done,
                                 double tol) {
  int not_in = 0;
#pragma omp parallel for reduction(+ : not_in)
  for (int s = 0; s < cap; ++s) {
    if (mask[s] > 0 && !done[s]) {
      int E = elem_ids[s];
      double p[3];
      load3(x_orig, stride, s, p);
      int ok;
      if (m->dim == 2) {
        int v[3]; double M[6], bcc[3];
        gather_tri(m, E, v, M);
        orc_barycentric_tri(m->vol[E], M, p, bcc);
        ok = orc_all_positive(bcc, 3, tol);
      } else {
        int v[4]; double M[12], bcc[4];
        gather_tet(m, E, v, M);
        orc_barycentric_tet(m->vol[E], M, p, bcc);
        ok = orc_all_positive(bcc, 4, tol);
      }
      if (!ok) {
        not_in += 1;
        elem_ids[s] = -1;
        done[s] = 1;
      }
    }
  }
  return not_in;
}

/* adjacency.tpp:232-364 find_exit_face */
static void find_exit_face(const orc_mesh* m, int cap, const unsigned char* mask,
                           const double* x_orig, const double* x_tgt, long stride,
                           int* elem_ids, int* done, int use_bcc, int* last_exit,
                           double* xpoints, double tol) {
  const int dim = m->dim;
  if (use_bcc && dim == 2) {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      if (mask[s] > 0 && !done[s]) {
        int E = elem_ids[s];
        int v[3]; double M[6], bcc[3], p[3];
        gather_tri(m, E, v, M);
        load3(x_tgt, stride, s, p);
        orc_barycentric_tri(m->vol[E], M, p, bcc);
        done[s] = orc_all_positive(bcc, 3, ORC_EPSILON);
        last_exit[s] = m->elem2sides[3 * (long)E + orc_min3(bcc)];
      }
    }
  } else if (use_bcc) {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      if (mask[s] > 0 && !done[s]) {
        int E = elem_ids[s];
        int v[4]; double M[12], bcc[4], p[3];
        gather_tet(m, E, v, M);
        load3(x_tgt, stride, s, p);
        /* :220-230 find_exit_face_bcc_3d */
        orc_barycentric_tet(m->vol[E], M, p, bcc);
        done[s] = orc_all_positive(bcc, 4, ORC_EPSILON);
        last_exit[s] = m->elem2sides[4 * (long)E + orc_min_index(bcc, 4)];
      }
    }
  } else if (dim == 2) {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      if (mask[s] > 0 && !done[s]) {
        int E = elem_ids[s];
        int tv[3]; double M[6], dest[3], orig[3];
        gather_tri(m, E, tv, M);
        load3(x_tgt, stride, s, dest);
        load3(x_orig, stride, s, orig);
        double xp[2] = {0, 0};
        const int prev = last_exit[s];
        last_exit[s] = -1;
        for (int ei = 0; ei < 3; ++ei) {
          int ed = m->elem2sides[3 * (long)E + ei];
          if (ed == prev) continue;
          int ev[2] = {m->side2verts[2 * (long)ed], m->side2verts[2 * (long)ed + 1]};
          double ec[4];
          for (int k = 0; k < 2; ++k)
            for (int i = 0; i < 2; ++i) ec[2 * k + i] = m->coords[2 * (long)ev[k] + i];
          int flip = orc_is_face_flipped_2d(ev, tv);
          if (orc_line_edge_2d(ec, orig, dest, xp, tol, flip)) {
            last_exit[s] = ed;
            xpoints[2 * (long)s] = xp[0];
            xpoints[2 * (long)s + 1] = xp[1];
          }
        }
        done[s] = (last_exit[s] == -1);
      }
    }
  } else {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      if (mask[s] > 0 && !done[s]) {
        int E = elem_ids[s];
        int tv[4]; double M[12], dest[3], orig[3];
        gather_tet(m, E, tv, M);
        load3(x_tgt, stride, s, dest);
        load3(x_orig, stride, s, orig);
        double xp[3] = {0, 0, 0};
        const int prev = last_exit[s];
        last_exit[s] = -1;
        double quality = -1;
        int best = -1;
        for (int fi = 0; fi < 4; ++fi) {
          int F = m->elem2sides[4 * (long)E + fi];
          if (F == prev) continue;
          int fv[3]; double fc[9];
          for (int k = 0; k < 3; ++k) {
            fv[k] = m->side2verts[3 * (long)F + k];
            for (int i = 0; i < 3; ++i) fc[3 * k + i] = m->coords[3 * (long)fv[k] + i];
          }
          int flip = orc_is_face_flipped_3d(fi, fv, tv);
          double dproj, closeness, param;
          int hit = orc_ray_intersects_triangle(fc, orig, dest, xp, tol, flip, &dproj,
                                                &closeness, &param);
          if (hit) {
            last_exit[s] = F;
            for (int i = 0; i < 3; ++i) xpoints[3 * (long)s + i] = xp[i];
          }
          if (dproj > -tol && (quality < 0 || closeness < quality) && last_exit[s] == -1) {
            quality = closeness;
            best = F;
            for (int i = 0; i < 3; ++i) xpoints[3 * (long)s + i] = xp[i];
          }
        }
        if (last_exit[s] == -1) last_exit[s] = best;
        done[s] = (last_exit[s] == -1);
      }
    }
  }
}

/* adjacency.tpp:366-387 check_model_intersection */
static void check_model_intersection(const orc_mesh* m, int cap, const unsigned char* mask,
                                     int* elem_ids, int* done, const int* last_exit,
                                     int require_intersection, int* xface) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s] > 0 && !done[s]) {
      const int bridge = last_exit[s];
      const int ex = m->exposed[bridge];
      done[s] = ex;
      if (ex && require_intersection) xface[s] = bridge;
      else elem_ids[s] = ex ? -1 : elem_ids[s];
    }
  }
}

/* adjacency.tpp:390-416 set_new_element */
static void set_new_element(const orc_mesh* m, int cap, const unsigned char* mask, int* elem_ids,
                            const int* done, const int* last_exit) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s] > 0 && !done[s]) {
      const int cur = elem_ids[s];
      const int first = m->side2elem_off[last_exit[s]];
      const int A = m->side2elem[first], B = m->side2elem[first + 1];
      elem_ids[s] = (A == cur) ? B : A;
    }
  }
}

int orc_search_mesh(const orc_mesh* m, int cap, const int* slot_elem, const unsigned char* mask,
                    const double* x_orig, const double* x_tgt, long stride, int* elem_ids,
                    int elem_ids_empty, int require_intersection, int* inter_faces,
                    double* inter_points, int inter_empty, int looplimit,
                    orc_search_stats* stats) {
  (void)inter_empty; /* fresh and reset arrays end up identical: zeros / -1 for every slot */
  const int dim = m->dim;
  int* done = (int*)calloc((size_t)cap + 1, sizeof(int));      /* :484 */
  int* last_exit = (int*)malloc(sizeof(int) * ((size_t)cap + 1)); /* :486 */
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) last_exit[s] = -1;
  const int use_bcc = !require_intersection;      /* :490 */
  const double tol = orc_compute_tolerance(m);    /* :491 */
  if (elem_ids_empty) {                           /* :504-515 */
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      elem_ids[s] = -1;
      if (mask[s]) elem_ids[s] = slot_elem[s];
      else done[s] = 1;
    }
  } else {                                        /* :516-522 */
#pragma omp parallel for
    for (int s = 0; s < cap; ++s)
      if ((mask[s] && elem_ids[s] == -1) || !mask[s]) done[s] = 1;
  }
  /* :525-533 finishUnmoved (3-component norm even in 2D) */
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s]) {
      double a[3], b[3], d[3];
      load3(x_orig, stride, s, a);
      load3(x_tgt, stride, s, b);
      v3_sub(b, a, d);
      if (v3_norm(d) < tol) done[s] = 1;
    }
  }
  if (require_intersection) {                     /* :535-549 */
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {
      for (int i = 0; i < dim; ++i) inter_points[(long)dim * s + i] = 0;
      inter_faces[s] = -1;
    }
  }
  int not_in = check_initial_parents(m, cap, mask, x_orig, stride, elem_ids, done, tol); /* :552 */
  int found = 0, loops = 0, not_found = 0;
  while (!found) {                                /* :558 */
    find_exit_face(m, cap, mask, x_orig, x_tgt, stride, elem_ids, done, use_bcc, last_exit,
                   inter_points, tol);
    check_model_intersection(m, cap, mask, elem_ids, done, last_exit, require_intersection,
                             inter_faces);
    set_new_element(m, cap, mask, elem_ids, done, last_exit);
    found = 1;
    if (min_done(done, cap) == 0) found = 0;      /* :568-572 */
    ++loops;
    if (looplimit && loops >= looplimit) {        /* :584-606 */
#pragma omp parallel for reduction(+ : not_found)
      for (int s = 0; s < cap; ++s)
        if (mask[s] > 0 && !done[s]) { elem_ids[s] = -1; not_found += 1; }
      break;
    }
  }
  if (stats) {
    stats->loops = loops; stats->not_in_elem = not_in; stats->not_found = not_found;
    stats->aborted = 0;
  }
  free(done); free(last_exit);
  return found;
}

/* ------------------------------------------------------------------------------------------
 * adjacency.hpp:1013-1158 search_mesh_2d
 * ---------------------------------------------------------------------------------------- */
int orc_search_mesh_2d(const orc_mesh* m, int cap, const int* slot_elem,
                       const unsigned char* mask, const double* x_tgt, long stride,
                       int* elem_ids, int looplimit, orc_search_stats* stats) {
  int* done = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  int* last_edge = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  const int nelems = m->nelems;
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {           /* :1045-1062 */
    done[s] = 1; last_edge[s] = -1;
    if (mask[s] > 0) {
      if (elem_ids[s] == -1) elem_ids[s] = slot_elem[s];
      done[s] = 0;
      if (elem_ids[s] == -nelems) { elem_ids[s] = -1; done[s] = 1; }
    } else {
      elem_ids[s] = -1; done[s] = 1;
    }
  }
  int found = 0, loops = 0, not_found = 0;
  while (!found) {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {          /* :1067-1084 checkCurrentElm */
      if (mask[s] > 0 && !done[s]) {
        int E = elem_ids[s];
        int v[3]; double M[6], bcc[3];
        gather_tri(m, E, v, M);
        double p[2] = {x_tgt[s], x_tgt[stride + s]};
        orc_barycentric_tri(m->vol[E], M, p, bcc);
        done[s] = orc_all_positive(bcc, 3, ORC_EPSILON);
        last_edge[s] = m->elem2sides[3 * (long)E + orc_min3(bcc)];
      }
    }
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {          /* :1086-1095 checkExposedEdges */
      if (mask[s] > 0 && !done[s]) {
        int ex = m->exposed[last_edge[s]];
        done[s] = ex;
        elem_ids[s] = ex ? -1 : elem_ids[s];
      }
    }
    set_new_element(m, cap, mask, elem_ids, done, last_edge); /* :1099-1117 */
    found = 1;
    if (min_done(done, cap) == 0) found = 0;
    ++loops;
    if (looplimit && loops >= looplimit) {   /* :1124-1147 */
#pragma omp parallel for reduction(+ : not_found)
      for (int s = 0; s < cap; ++s)
        if (mask[s] > 0 && !done[s]) { elem_ids[s] = -1; not_found += 1; }
      break;
    }
  }
  if (stats) {
    stats->loops = loops; stats->not_in_elem = 0; stats->not_found = not_found;
    stats->aborted = 0;
  }
  free(done); free(last_edge);
  return found;
}

/* ------------------------------------------------------------------------------------------
 * adjacency.hpp:559-768 legacy 3D search_mesh
 * ---------------------------------------------------------------------------------------- */
int orc_search_mesh_legacy3d(const orc_mesh* m, int cap, const int* slot_elem,
                             const unsigned char* mask, const double* x_orig,
                             const double* x_tgt, long stride, int* elem_ids,
                             int elem_ids_empty, double* xpoints_d, int* xface_d,
                             int looplimit, orc_search_stats* stats) {
  const double tol = 1.0e-10;
  int* done = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  int* next = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  const int ndual = m->dual_off[m->nelems];
  int aborted = 0;
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {           /* :586-598 fill */
    next[s] = -1;
    if (mask[s] > 0) {
      if (elem_ids_empty) elem_ids[s] = slot_elem[s];
      done[s] = (elem_ids[s] == -1);
    } else {
      elem_ids[s] = -1; done[s] = 1;
    }
  }
  int found = 0, loops = 0;
  while (!found) {
#pragma omp parallel for reduction(+ : aborted)
    for (int s = 0; s < cap; ++s) {          /* :607-740 adj_search */
      if (!(mask[s] > 0 && !done[s])) continue;
      const int E = elem_ids[s];
      int tv[4]; double M[12], dest[3], orig[3], bcc[4];
      gather_tet(m, E, tv, M);
      load3(x_tgt, stride, s, dest);
      load3(x_orig, stride, s, orig);
      if (loops == 0) {
        orc_find_barycentric_tet(M, orig, bcc);
        if (!orc_all_positive(bcc, 4, tol)) aborted += 1; /* OMEGA_H_CHECK(false) :626 */
      }
      int intersected = 0;
      orc_find_barycentric_tet(M, dest, bcc);
      if (orc_all_positive(bcc, 4, tol)) {
        next[s] = E;
        done[s] = 1;
        continue;
      }
      double dproj[4] = {-1, -1, -1, -1};
      double xpts[12] = {0};
      int exposed_faces[4], xface_ids[4];
      int dual_id = m->dual_off[E];
      int findex = 0;
      for (int iface = 0; iface < 4; ++iface) {
        const int F = m->elem2sides[4 * (long)E + iface];
        double xp[3];
        const int ex = m->exposed[F];
        exposed_faces[findex] = ex;
        xface_ids[findex] = F;
        int fv[3]; double fc[9];
        for (int k = 0; k < 3; ++k) {
          fv[k] = m->side2verts[3 * (long)F + k];
          for (int i = 0; i < 3; ++i) fc[3 * k + i] = m->coords[3 * (long)fv[k] + i];
        }
        const int m1 = FACE_MAP[findex * 2], m2 = FACE_MAP[findex * 2 + 1];
        int flip = 1;
        if (fv[1] == tv[m1] && fv[2] == tv[m2]) flip = 0;
        intersected = orc_line_triangle_intx_simple(fc, orig, dest, xp, &dproj[findex], flip, tol);
        for (int i = 0; i < 3; ++i) xpts[findex * 3 + i] = xp[i];
        if (intersected && ex) {
          done[s] = 1;
          for (int i = 0; i < 3; ++i) xpoints_d[3 * (long)s + i] = xp[i];
          xface_d[s] = F;
          next[s] = -1;
          break;
        } else if (intersected && !ex) {
          next[s] = m->dual[dual_id];
          break;
        }
        if (!ex) ++dual_id;
        ++findex;
      }
      if (!intersected) {                    /* :714-738 */
        const int mi = orc_max_index(dproj, 4);
        if (dproj[mi] >= 0) {
          const int fid = xface_ids[mi];
          if (exposed_faces[mi]) {
            next[s] = -1;
            for (int i = 0; i < 3; ++i) xpoints_d[3 * (long)s + i] = xpts[mi * 3 + i];
            xface_d[s] = fid;
            done[s] = 1;
          } else {
            /* reference indexes the dual graph by FACE id here (:726); reproduced.  An
             * out-of-range read is undefined in the reference; the oracle drops the particle. */
            if (fid < ndual) next[s] = m->dual[fid];
            else { next[s] = -1; done[s] = 1; }
          }
        } else {
          next[s] = -1;
          done[s] = 1;
        }
      }
    }
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) elem_ids[s] = next[s];  /* :745-748 */
    found = 1;
    if (min_done(done, cap) == 0) found = 0;
    ++loops;
    if (looplimit && loops > looplimit) break;            /* :756 (note: '>') */
  }
  if (stats) {
    stats->loops = loops; stats->not_in_elem = 0; stats->not_found = 0;
    stats->aborted = aborted;
  }
  free(done); free(next);
  return found;
}

/* ------------------------------------------------------------------------------------------
 * adjacency.hpp:316-555 search_mesh_3d (three kernels per iteration: checkCurrentElm,
 * findIntersection, processUndetected), with barycentric_coords_tet :136-158 and
 * isPointWithinElemTet :300-313, tol 1e-20.
 * ---------------------------------------------------------------------------------------- */
/* adjacency.hpp:136-158: vals = 1/6 (p-a).cross(c-a, b-a), vol = tet_volume_from_basis; when
 * vol < tol the function returns with bcc zeroed (the caller ignores the return value). */
int orc_barycentric_coords_tet(const double M[12], const double p[3], double bcc[4], double tol) {
  double vals[4];
  for (int f = 0; f < 4; ++f) {
    const double* a = M + 3 * TET_FACE[f][0];
    const double* b = M + 3 * TET_FACE[f][1];
    const double* c = M + 3 * TET_FACE[f][2];
    double vab[3], vac[3], vap[3], cr[3];
    v3_sub(b, a, vab);
    v3_sub(c, a, vac);
    v3_sub(p, a, vap);
    v3_cross(vac, vab, cr);
    vals[f] = 1.0 / 6.0 * v3_dot(vap, cr);
    bcc[f] = 0;
  }
  double b0[3], b1[3], b2[3], c[3];
  v3_sub(M + 3, M, b0);
  v3_sub(M + 6, M, b1);
  v3_sub(M + 9, M, b2);
  v3_cross(b0, b1, c);
  const double vol = v3_dot(c, b2) / 6.0;
  if (vol < tol) return 0;
  const double inv = 1.0 / vol;
  for (int i = 0; i < 4; ++i) bcc[i] = inv * vals[i];
  return 1;
}

static int point_within_tet(const orc_mesh* m, int e, const double p[3], double tol) {
  int tv[4]; double M[12], bcc[4];
  gather_tet(m, e, tv, M);
  orc_barycentric_coords_tet(M, p, bcc, tol);
  return orc_all_positive(bcc, 4, tol);
}

int orc_search_mesh_3d(const orc_mesh* m, int cap, const int* slot_elem,
                       const unsigned char* mask, const double* x_orig, const double* x_tgt,
                       long stride, int* elem_ids, int elem_ids_empty, double* xpoints_d,
                       int* xface_d, int looplimit, orc_search_stats* stats) {
  const double tol = 1.0e-20;
  int* done = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  int* next = (int*)malloc(sizeof(int) * ((size_t)cap + 1));
  const int ndual = m->dual_off[m->nelems];
  int aborted = 0;
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {            /* :356-366 fill */
    next[s] = -1;
    if (mask[s] > 0) {
      if (elem_ids_empty) elem_ids[s] = slot_elem[s];
      done[s] = (elem_ids[s] == -1) * 2;
    } else {
      elem_ids[s] = -1; done[s] = 2;
    }
  }
#pragma omp parallel for reduction(+ : aborted)
  for (int s = 0; s < cap; ++s) {            /* :368-379 checkParent: tests the ROW element */
    if (!(mask[s] > 0 && done[s] != 2)) continue;
    double orig[3];
    load3(x_orig, stride, s, orig);
    if (!point_within_tet(m, slot_elem[s], orig, tol)) aborted += 1;   /* OMEGA_H_CHECK(false) */
  }
  int found = 0, loops = 0;
  while (!found) {
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {          /* :395-410 checkCurrentElm */
      if (!(mask[s] > 0 && !done[s])) continue;
      double dest[3];
      load3(x_tgt, stride, s, dest);
      done[s] = point_within_tet(m, elem_ids[s], dest, tol) ? 2 : 0;
      next[s] = elem_ids[s];
    }
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {          /* :412-469 findIntersection */
      if (!(mask[s] > 0 && done[s] < 2)) continue;
      const int E = elem_ids[s];
      int tv[4]; double M[12], dest[3], orig[3];
      gather_tet(m, E, tv, M);
      load3(x_tgt, stride, s, dest);
      load3(x_orig, stride, s, orig);
      int dual_id = m->dual_off[E];
      int adj_id = -1, ind_exp = -1;
      double projd[4] = {0, 0, 0, 0};
      double xpts[3] = {0, 0, 0};
      for (int fi = 0; fi < 4; ++fi) {
        const int F = m->elem2sides[4 * (long)E + fi];
        int fv[3]; double fc[9], xp[3];
        for (int k = 0; k < 3; ++k) {
          fv[k] = m->side2verts[3 * (long)F + k];
          for (int i = 0; i < 3; ++i) fc[3 * k + i] = m->coords[3 * (long)fv[k] + i];
        }
        const int flip = orc_is_face_flipped_3d(fi, fv, tv);
        const int det = orc_line_triangle_intx_simple(fc, orig, dest, xp, &projd[fi], flip, tol);
        const int ex = m->exposed[F];
        if (det && ex) {
          ind_exp = fi;
          for (int i = 0; i < 3; ++i) xpts[i] = xp[i];
        }
        if (det && !ex) adj_id = dual_id;
        if (!ex) ++dual_id;
      }
      if (ind_exp >= 0) {                    /* wall collision */
        for (int i = 0; i < 3; ++i) xpoints_d[3 * (long)s + i] = xpts[i];
        xface_d[s] = m->elem2sides[4 * (long)E + ind_exp];
        next[s] = -1;
        done[s] = 2;
      }
      if (adj_id >= 0) {                     /* interior; also overrides a wall hit (:460-468) */
        next[s] = m->dual[adj_id];
        done[s] = 1;
      }
    }
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) {          /* :471-516 processUndetected */
      const int d = done[s];
      done[s] = (d < 2) ? 0 : 2;
      if (!(mask[s] > 0 && d < 1)) continue;
      const int E = elem_ids[s];
      int tv[4]; double M[12], dest[3], orig[3];
      gather_tet(m, E, tv, M);
      load3(x_tgt, stride, s, dest);
      load3(x_orig, stride, s, orig);
      double projd[4] = {-1, -1, -1, -1};
      double xpoints[12];
      for (int fi = 0; fi < 4; ++fi) {
        const int F = m->elem2sides[4 * (long)E + fi];
        int fv[3]; double fc[9], xp[3];
        for (int k = 0; k < 3; ++k) {
          fv[k] = m->side2verts[3 * (long)F + k];
          for (int i = 0; i < 3; ++i) fc[3 * k + i] = m->coords[3 * (long)fv[k] + i];
        }
        const int flip = orc_is_face_flipped_3d(fi, fv, tv);
        orc_line_triangle_intx_simple(fc, orig, dest, xp, &projd[fi], flip, tol);
        for (int i = 0; i < 3; ++i) xpoints[fi * 3 + i] = xp[i];
      }
      const int mi = orc_max_index(projd, 4);
      const int fid = m->elem2sides[4 * (long)E + mi];
      if (m->exposed[fid]) {
        next[s] = -1;
        for (int i = 0; i < 3; ++i) xpoints_d[3 * (long)s + i] = xpoints[mi * 3 + i];
        xface_d[s] = fid;
        done[s] = 2;
      } else {
        /* dual graph indexed by FACE id (:510), as in the legacy search; an out-of-range read
         * is undefined in the reference, the oracle drops the particle */
        if (fid < ndual) next[s] = m->dual[fid];
        else { next[s] = -1; done[s] = 2; }
      }
    }
#pragma omp parallel for
    for (int s = 0; s < cap; ++s) elem_ids[s] = next[s];  /* :518-522 */
    found = 1;
    if (min_done(done, cap) == 0) found = 0;
    ++loops;
    if (looplimit && loops >= looplimit) break;           /* :528 */
  }
  if (stats) {
    int nf = 0;
    for (int s = 0; s < cap; ++s) nf += (mask[s] > 0 && !done[s]);
    stats->loops = loops; stats->not_in_elem = 0; stats->not_found = nf;
    stats->aborted = aborted;
  }
  free(done); free(next);
  return found;
}

/* ------------------------------------------------------------------------------------------
 * Gather (mesh / grid -> particle field interpolation): adjacency.hpp:770-809 and
 * pumipic_utils.hpp:186-456.  Device helpers of the reference that GITRm's Boris push calls.
 * ---------------------------------------------------------------------------------------- */
/* adjacency.hpp:772-790 interpolateTetVtx.  The reference indexes its 4-entry gather with
 * d*dof+comp, which is only in bounds for dof == 1; the restatement reads field[vert*dof+comp]
 * (the documented intent: "Field has dof components ... stored in order 0,1,2,3 at tet's
 * vertices"), identical to the reference for dof == 1. */
double orc_interpolate_tet_vtx(const orc_mesh* m, const double* field, int elem,
                               const double bcc[4], int dof, int comp) {
  static const int OPP[4] = {3, 2, 0, 1};   /* simplex_opposite_template(3, 2, fi) */
  const int* tv = m->elem2verts + 4 * (long)elem;
  double val = 0;
  for (int fi = 0; fi < 4; ++fi) val = val + bcc[fi] * field[(long)tv[OPP[fi]] * dof + comp];
  return val;
}
/* adjacency.hpp:801-809 findBCCoordsInTet + :793-799 interpolate3dFieldTet for every masked
 * particle: out[c*stride+s].  Returns the number of particles the reference would abort on
 * (find_barycentric_tet failed or a coordinate below -EPSILON); their output is left untouched. */
int orc_gather_tet_field(const orc_mesh* m, int cap, const unsigned char* mask, const double* x,
                         long stride, const int* elem_ids, const double* field, int dof,
                         double* out) {
  int bad = 0;
#pragma omp parallel for reduction(+ : bad)
  for (int s = 0; s < cap; ++s) {
    if (!mask[s]) continue;
    const int e = elem_ids[s];
    if (e < 0) continue;
    int tv[4]; double M[12], p[3], bcc[4];
    gather_tet(m, e, tv, M);
    load3(x, stride, s, p);
    const int res = orc_find_barycentric_tet(M, p, bcc);
    if (!res || !orc_all_positive(bcc, 4, 1e-10)) { bad += 1; continue; }
    for (int c = 0; c < dof; ++c) out[c * stride + s] = orc_interpolate_tet_vtx(m, field, e, bcc, dof, c);
  }
  return bad;
}

/* pumipic_utils.hpp:245-248 */
static inline double interp2d_base(double d1, double d2, double grid1, double grid2, double v,
                                   double dv) {
  return (d1 * (grid2 - v) + d2 * (v - grid1)) / dv;
}
/* pumipic_utils.hpp:260-296 interpolate2d */
static double interpolate2d(const double* data, double gridXi, double gridXip1, double gridZj,
                            double gridZjp1, double x0, double z, int nx, int nz, int i, int j,
                            double dx, double dz, double y, int cyl, int nComp, int comp) {
  if (nx <= 1 && nz <= 1) return data[comp];
  double x = x0;
  if (cyl) x = sqrt(x * x + y * y);
  double fxz = 0;
  if (i >= nx - 1 && j >= nz - 1) {
    fxz = data[(nx - 1 + (long)(nz - 1) * nx) * nComp + comp];
  } else if (i >= nx - 1) {
    fxz = interp2d_base(data[(nx - 1 + (long)j * nx) * nComp + comp],
                        data[(nx - 1 + (long)(j + 1) * nx) * nComp + comp], z - gridZj, gridZjp1 - z,
                        z, dz);
  } else if (j >= nz - 1) {
    fxz = interp2d_base(data[(i + (long)(nz - 1) * nx) * nComp + comp],
                        data[(i + (long)(nz - 1) * nx) * nComp + comp], x - gridXi, gridXip1 - x, x,
                        dx);
  } else {
    const double f1 = interp2d_base(data[(i + (long)j * nx) * nComp + comp],
                                    data[(i + 1 + (long)j * nx) * nComp + comp], gridXi, gridXip1, x, dx);
    const double f2 = interp2d_base(data[(i + (long)(j + 1) * nx) * nComp + comp],
                                    data[(i + 1 + (long)(j + 1) * nx) * nComp + comp], gridXi, gridXip1,
                                    x, dx);
    fxz = interp2d_base(f1, f2, gridZj, gridZjp1, z, dz);
  }
  return fxz;
}
/* pumipic_utils.hpp:298-321 interpolate2d_field (uniform grid given by origin and spacing) */
double orc_interpolate2d_field(const double* data, double gridx0, double gridz0, double dx,
                               double dz, int nx, int nz, const double pos[3], int cyl, int nComp,
                               int comp) {
  if (nx <= 1 && nz <= 1) return data[comp];
  double x = pos[0];
  const double z = pos[2];
  if (cyl) x = sqrt(x * x + pos[1] * pos[1]);
  int i = (int)floor((x - gridx0) / dx);
  int j = (int)floor((z - gridz0) / dz);
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  const double gridXi = gridx0 + i * dx, gridXip1 = gridx0 + (i + 1) * dx;
  const double gridZj = gridz0 + j * dz, gridZjp1 = gridz0 + (j + 1) * dz;
  return interpolate2d(data, gridXi, gridXip1, gridZj, gridZjp1, x, z, nx, nz, i, j, dx, dz, 0, 0,
                       nComp, comp);
}
/* pumipic_utils.hpp:439-456 interp2dVector: three components, rotated by atan2(y, x) when the
 * data are cylindrically symmetric */
void orc_interp2d_vector(const double* data3, double gridx0, double gridz0, double dx, double dz,
                         int nx, int nz, const double pos[3], double field[3], int cyl) {
  for (int i = 0; i < 3; ++i)
    field[i] = orc_interpolate2d_field(data3, gridx0, gridz0, dx, dz, nx, nz, pos, cyl, 3, i);
  if (cyl) {
    const double theta = atan2(pos[1], pos[0]);
    const double f0 = field[0], f1 = field[1];
    field[0] = cos(theta) * f0 - sin(theta) * f1;
    field[1] = sin(theta) * f0 + cos(theta) * f1;
  }
}
/* pumipic_utils.hpp:377-420 interpolate3d_field (grid coordinates given as arrays) */
double orc_interpolate3d_field(double x, double y, double z, int nx, int ny, int nz,
                               const double* gridx, const double* gridy, const double* gridz,
                               const double* data) {
  /* The reference evaluates all four rows and the y / z spacings unconditionally and discards
   * them when ny <= 1 or nz <= 1 (:415-416); those reads lie outside the tables, so the
   * restatement skips what is discarded. */
  const double dx = gridx[1] - gridx[0];
  const double dy = ny > 1 ? gridy[1] - gridy[0] : 1.0;
  const double dz = nz > 1 ? gridz[1] - gridz[0] : 1.0;
  int i = (int)floor((x - gridx[0]) / dx);
  int j = (int)floor((y - gridy[0]) / dy);
  int k = (int)floor((z - gridz[0]) / dz);
  i = (i < 0) ? 0 : ((i >= nx - 1) ? (nx - 2) : i);
  j = (j < 0 || ny <= 1) ? 0 : ((j >= ny - 1) ? (ny - 2) : j);
  k = (k < 0 || nz <= 1) ? 0 : ((k >= nz - 1) ? (nz - 2) : k);
#define D_(I) data[(I)], data[(I) + 1]
  const long nxy = (long)nx * ny;
  const double fx_z0 = interp2d_base(D_(i + (long)j * nx + k * nxy), gridx[i], gridx[i + 1], x, dx);
  if (nz <= 1) return fx_z0;
  const double fx_z1 = interp2d_base(D_(i + (long)j * nx + (k + 1) * nxy), gridx[i], gridx[i + 1], x, dx);
  const double fxz0 = interp2d_base(fx_z0, fx_z1, gridz[k], gridz[k + 1], z, dz);
  if (ny <= 1) return fxz0;
  const double fxy_z0 = interp2d_base(D_(i + (long)(j + 1) * nx + k * nxy), gridx[i], gridx[i + 1], x, dx);
  const double fxy_z1 = interp2d_base(D_(i + (long)(j + 1) * nx + (k + 1) * nxy), gridx[i], gridx[i + 1], x, dx);
#undef D_
  const double fxz1 = interp2d_base(fxy_z0, fxy_z1, gridz[k], gridz[k + 1], z, dz);
  return interp2d_base(fxz0, fxz1, gridy[j], gridy[j + 1], y, dy);
}
/* per-particle drivers over a particle structure (masked slots) */
void orc_gather_grid2d_vector(int cap, const unsigned char* mask, const double* x, long stride,
                              const double* data3, double gridx0, double gridz0, double dx,
                              double dz, int nx, int nz, int cyl, double* out) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (!mask[s]) continue;
    double p[3], f[3];
    load3(x, stride, s, p);
    orc_interp2d_vector(data3, gridx0, gridz0, dx, dz, nx, nz, p, f, cyl);
    for (int c = 0; c < 3; ++c) out[c * stride + s] = f[c];
  }
}
void orc_gather_grid3d(int cap, const unsigned char* mask, const double* x, long stride,
                       const double* data, const double* gridx, const double* gridy,
                       const double* gridz, int nx, int ny, int nz, double* out) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (!mask[s]) continue;
    out[s] = orc_interpolate3d_field(x[s], x[stride + s], x[2 * stride + s], nx, ny, nz, gridx,
                                     gridy, gridz, data);
  }
}

/* ------------------------------------------------------------------------------------------
 * Pushes
 * ---------------------------------------------------------------------------------------- */
void orc_push_constant(int cap, const unsigned char* mask, const double* x, double* xtgt,
                       long stride, double distance, double dx, double dy, double dz) {
  const double disp[4] = {distance, dx, dy, dz};
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s]) {
      const double unique = 0.0; /* ptclUnique_d is a zero-filled array (:100) */
      for (int i = 0; i < 3; ++i) {
        double dir = disp[0] * disp[i + 1];
        xtgt[i * stride + s] = x[i * stride + s] + dir + unique;
      }
    }
  }
}

void orc_push_direction(int cap, const unsigned char* mask, double* tgt, const double* dir,
                        long stride, double distance) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s)
    if (mask[s])
      for (int i = 0; i < 3; ++i)
        tgt[i * stride + s] = tgt[i * stride + s] + distance * dir[i * stride + s];
}

void orc_elliptical_setup(int cap, const unsigned char* mask, const double* x, long stride,
                          float* b, float* phi, double h, double k, double d) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s]) {
      const double w = x[s], z = x[stride + s];
      const double ph = atan2(d * (z - k), w - h);
      const double bb = (z - k) / sin(ph);
      phi[s] = (float)ph;
      b[s] = (float)bb;
    }
  }
}

void orc_elliptical_push(int cap, const int* slot_elem, const unsigned char* mask, double* xtgt,
                         long stride, const float* b, float* phi, const int* class_ids,
                         double h, double k, double d, double deg) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    if (mask[s]) {
      const int cls = class_ids[slot_elem[s]];
      const double centerFactor = cls == 1 ? 0.01 : 1.0;
      const double distByClass = centerFactor * (double)1.0 / cls;
      const double degP = deg * distByClass;
      const float ph = phi[s];
      const float bb = b[s];
      const double a = bb * d;
      const double rad = ph + degP * M_PI / 180.0;
      xtgt[s] = a * cos(rad) + h;
      xtgt[stride + s] = bb * sin(rad) + k;
      phi[s] = (float)rad;
    }
  }
}

/* n particles, AoS-free flat arrays pos[3][n] etc. (component-major, stride n) */
void orc_push_boris(int n, double* pos, double* pos_prev, double* vel, const double* efield,
                    const double* bfield, double dt) {
#pragma omp parallel for
  for (int p = 0; p < n; ++p) {
    double v[3], E[3], B[3];
    for (int i = 0; i < 3; ++i) {
      v[i] = vel[(long)i * n + p]; E[i] = efield[(long)i * n + p]; B[i] = bfield[(long)i * n + p];
    }
    const double charge = 1, amu = 10;
    const double bmag = v3_norm(B);
    const double qPrime = charge * 1.60217662e-19 / (amu * 1.6737236e-27) * dt * 0.5;
    const double coeff = 2.0 * qPrime / (1.0 + (qPrime * bmag) * (qPrime * bmag));
    double qpE[3], vMinus[3], c1[3], vPrime[3], c2[3];
    for (int i = 0; i < 3; ++i) qpE[i] = E[i] * qPrime;
    v3_sub(v, qpE, vMinus);
    v3_cross(vMinus, B, c1);
    for (int i = 0; i < 3; ++i) vPrime[i] = vMinus[i] + c1[i] * qPrime;
    v3_cross(vPrime, B, c2);
    for (int i = 0; i < 3; ++i) v[i] = vMinus[i] + c2[i] * coeff;
    for (int i = 0; i < 3; ++i) v[i] = v[i] + qpE[i];
    for (int i = 0; i < 3; ++i) {
      const double pre = pos_prev[(long)i * n + p];
      pos_prev[(long)i * n + p] = pos[(long)i * n + p];
      pos[(long)i * n + p] = pre + v[i] * dt;
      vel[(long)i * n + p] = v[i];
    }
  }
}

void orc_update_positions(int cap, double* x, double* xtgt, long stride) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s)
    for (int i = 0; i < 3; ++i) {
      x[i * stride + s] = xtgt[i * stride + s];
      xtgt[i * stride + s] = 0;
    }
}

void orc_set_unsafe_procs(int cap, const unsigned char* mask, const int* elems, const int* safe,
                          const int* owner, int self, int* new_elems, int* new_procs) {
#pragma omp parallel for
  for (int s = 0; s < cap; ++s) {
    new_procs[s] = self;
    const int nelm = elems[s];
    new_elems[s] = nelm;
    if (mask[s] && nelm != -1 && !safe[nelm]) new_procs[s] = owner[nelm];
  }
}

/* ------------------------------------------------------------------------------------------
 * Gyro-averaged scatter: test/gyroScatter.hpp
 * ---------------------------------------------------------------------------------------- */
void orc_gyro_scatter(const orc_mesh* m, int cap, const int* slot_elem,
                      const unsigned char* mask, const int* v2v, double rmax, int gnr, int gppr,
                      double* scatter_w) {
  const int nverts = m->nverts;
  const double ringWidth = rmax / gnr;
  double* ring_accum = (double*)calloc((size_t)gnr * nverts, sizeof(double));
  /* :182-204 accumulateToRings -- serial on purpose: the reference's atomic order is
   * arbitrary and every addend is 1.0, so any order gives the same exact integers. */
  for (int s = 0; s < cap; ++s) {
    if (mask[s] > 0) {
      const double ptclRadius = ringWidth * 1.125;
      int ringDown = 0;
      for (int i = 2; i <= gnr; i++) ringDown += (ptclRadius >= ringWidth * i);
      const int ringUp = ringDown + 1;
      const int e = slot_elem[s];
      for (int i = 0; i < 3; i++) {
        const int v = m->elem2verts[3 * (long)e + i];
        ring_accum[(long)v * gnr + ringUp] += 1;
        ring_accum[(long)v * gnr + ringDown] += 1;
      }
    }
  }
  for (int v = 0; v < nverts; ++v) scatter_w[v] = 0;
  /* :207-224 scatterToMappedVerts (vertex order ascending = one legal atomic order) */
  for (int v = 0; v < nverts; ++v) {
    const long vtxIdx = (long)v * gnr * gppr;
    for (int ring = 0; ring < gnr; ring++) {
      const double val = ring_accum[(long)v * gnr + ring] / gppr;
      for (int pt = 0; pt < gppr; pt++) {
        const long ptIdx = 3 * (vtxIdx + (long)ring * gppr + pt);
        for (int k = 0; k < 3; k++) {
          const int mv = v2v[ptIdx + k];
          if (mv >= 0) scatter_w[mv] += val;
        }
      }
    }
  }
  free(ring_accum);
}

int orc_gyro_ring_map(const orc_mesh* m, double rmax, int gnr, int gppr, double theta_deg,
                      int* map) {
  const int nverts = m->nverts;
  const long npts = (long)nverts * gnr * gppr;
  const double torad = M_PI / 180;
  /* the throw-away particle structure of :45-52 is replaced by a flat slot list: slot == point */
  double* tgt = (double*)calloc((size_t)npts * 3, sizeof(double));
  int* start = (int*)malloc(sizeof(int) * (size_t)npts);
  unsigned char* mask = (unsigned char*)malloc((size_t)npts);
  int* elem_ids = (int*)malloc(sizeof(int) * (size_t)npts);
  for (long id = 0; id < npts; ++id) {      /* :111-121 generateRingPoints */
    const int point_id = (int)(id % gppr);
    const long id2 = id / gppr;
    const int ring_id = (int)(id2 % gnr);
    const int vert_id = (int)(id2 / gnr);
    const double radius = rmax * (ring_id + 1) / gnr;
    const double deg = theta_deg + (((double)point_id) / gppr * 360);
    const double rad = deg * torad;
    tgt[id] = m->coords[2 * (long)vert_id] + radius * cos(rad);
    tgt[npts + id] = m->coords[2 * (long)vert_id + 1] + radius * sin(rad);
    start[id] = m->vert2elem[m->vert2elem_off[vert_id]]; /* :138-144 first adjacent element */
    mask[id] = 1;
    elem_ids[id] = -1;
  }
  orc_search_stats st;
  int found = orc_search_mesh_2d(m, (int)npts, start, mask, tgt, npts, elem_ids, 100, &st);
  for (long id = 0; id < npts; ++id) {      /* :73-86 createGyroMapping */
    const int parent = elem_ids[id];
    for (int i = 0; i < 3; ++i)
      map[3 * id + i] = parent >= 0 ? m->elem2verts[3 * (long)parent + i] : -1;
  }
  free(tgt); free(start); free(mask); free(elem_ids);
  return found;
}
