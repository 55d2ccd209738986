// pp_search.cu -- particle-to-element adjacency search, one fused kernel per call.
//
// Replaces (all in src/): pumipic_adjacency.tpp search_mesh :642 / trace_particle_through_mesh
// :461 (setInitial :504-522, finishUnmoved :525-533, initializeIntersection :535-549,
// check_initial_parents :73-145, find_exit_face :232-364, check_model_intersection :366-387,
// set_new_element :390-416, the per-iteration get_min :568-572 and the loop-limit sweep
// :584-606); pumipic_adjacency.hpp search_mesh_2d :1013-1158 and the legacy 3D search_mesh
// :559-768.  The reference launches 3 kernels + 1 device->host reduction per walk iteration
// over the whole capacity; here each thread owns one slot and walks to completion, every hop
// being ONE aligned gather of a packed walk record (pp_internal.cuh).  Arithmetic follows the
// reference operation by operation (the library is built with -fmad=false) so element ids are
// bit-identical to the CPU oracle.
#include <type_traits>

#include "pp_internal.cuh"

namespace {

enum Mode { M_BCC = 0, M_RAY = 1, M_LEG2D = 2, M_LEG3D = 3, M_S3D = 4 };

struct SearchParams {
  PsView ps;
  const void* walk;
  const void* walk_bcc;
  int staged;           // 1: block-staged BCC walk (default), 0: thread-per-slot reference kernel
  const double* xo;
  const double* xt;
  long stride;
  int* elem_ids;
  int ids_empty;
  int* inter_faces;
  double* inter_points;
  int looplimit;
  double tol;
  double unmoved_sq;    // see unmoved_threshold()
  int chunk_begin, chunk_end;   // chunk walk: chunks [begin, end) of the structure
  int* sched;           // chunk walk: dynamic scheduler counter of this launch
  int nelems;
  SearchCounters* counters;
  // fused direction push (test_adj.cpp:550-562): xt += distance*dir before the walk
  const double* dir;
  double distance;
  double* xt_rw;
  int push_from_orig;   // 1: xt = xo + distance*dir (PIC form), 0: xt += distance*dir (test_adj)
  // legacy 3D fallback (adjacency.hpp:726 indexes the dual graph by face id)
  const int* elem2sides;
  const int* dual;
  int ndual;
};

constexpr double kEps = 1e-10;  // pumipic_constants.hpp:6

// ---------------------------------------------------------------- record loads
struct Tet {
  d3 M[4];
  double vol;
  int adj[4];
  unsigned codes;
  int aux;
};
struct Tri {
  d2 M[3];
  double area;
  int adj[3];
  unsigned codes;
  int cls, aux;
};

template <class R>
__device__ __forceinline__ int adj_of(const R& r, int f) {   // no dynamic register indexing
#ifdef PP_AB_OLD_ADJ
  return r.adj[f];
#else
  return f == 0 ? r.adj[0] : f == 1 ? r.adj[1] : f == 2 ? r.adj[2] : r.adj[3];
#endif
}
__device__ __forceinline__ int adj_of(const Tri& r, int f) {
  return f == 0 ? r.adj[0] : f == 1 ? r.adj[1] : r.adj[2];
}
__device__ __forceinline__ d3 vert_of(const Tet& t, unsigned i) {
  return i == 0 ? t.M[0] : i == 1 ? t.M[1] : i == 2 ? t.M[2] : t.M[3];
}
__device__ __forceinline__ d2 vert_of(const Tri& t, unsigned i) {
  return i == 0 ? t.M[0] : i == 1 ? t.M[1] : t.M[2];
}

__device__ __forceinline__ void load_rec(const void* walk, int E, Tet& t) {
  const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const PPTetRec*>(walk) + E);
  const double2 a0 = __ldg(p + 0), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
  const double2 a3 = __ldg(p + 3), a4 = __ldg(p + 4), a5 = __ldg(p + 5);
  const int4 q6 = __ldg(reinterpret_cast<const int4*>(p + 6));
  const int4 q7 = __ldg(reinterpret_cast<const int4*>(p + 7));
  t.M[0] = {a0.x, a0.y, a1.x};
  t.M[1] = {a1.y, a2.x, a2.y};
  t.M[2] = {a3.x, a3.y, a4.x};
  t.M[3] = {a4.y, a5.x, a5.y};
  t.vol = __hiloint2double(q6.y, q6.x);
  t.adj[0] = q6.z; t.adj[1] = q6.w; t.adj[2] = q7.x; t.adj[3] = q7.y;
  t.codes = (unsigned)q7.z;
  t.aux = q7.w;
}
__device__ __forceinline__ void load_rec(const void* walk, int E, Tri& t) {
  const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const PPTriRec*>(walk) + E);
  const double2 a0 = __ldg(p + 0), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
  const int4 q3 = __ldg(reinterpret_cast<const int4*>(p + 3));
  const int4 q4 = __ldg(reinterpret_cast<const int4*>(p + 4));
  t.M[0] = {a0.x, a0.y};
  t.M[1] = {a1.x, a1.y};
  t.M[2] = {a2.x, a2.y};
  t.area = __hiloint2double(q3.y, q3.x);
  t.adj[0] = q3.z; t.adj[1] = q3.w; t.adj[2] = q4.x;
  t.codes = (unsigned)q4.y;
  t.cls = q4.z;
  t.aux = q4.w;
}

// ---------------------------------------------------------------- barycentric coordinates
// adjacency.tpp:41-69 barycentric_tet.  Faces (0,2,1),(0,1,3),(1,2,3),(2,0,3):
// vals[f] = (p - a) . cross(c - a, b - a);  bcc = (1/vol) * vals  (sums to 6)
__device__ __forceinline__ void bcc_tet(const Tet& t, d3 p, double bcc[4]) {
  const d3 n0 = cross3(t.M[1] - t.M[0], t.M[2] - t.M[0]);
  const d3 n1 = cross3(t.M[3] - t.M[0], t.M[1] - t.M[0]);
  const d3 n2 = cross3(t.M[3] - t.M[1], t.M[2] - t.M[1]);
  const d3 n3 = cross3(t.M[3] - t.M[2], t.M[0] - t.M[2]);
  const d3 p0 = p - t.M[0];
  double v[4];
  v[0] = dot3(p0, n0);
  v[1] = dot3(p0, n1);
  v[2] = dot3(p - t.M[1], n2);
  v[3] = dot3(p - t.M[2], n3);
  if (t.vol > 0) {
    const double inv = 1.0 / t.vol;
#pragma unroll
    for (int i = 0; i < 4; ++i) bcc[i] = inv * v[i];
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) bcc[i] = -1;
  }
}
// adjacency.hpp:97-133 find_barycentric_tet: same vals, own vol6 (> 1e-20), sums to 1
__device__ __forceinline__ bool bcc_tet_legacy(const Tet& t, d3 p, double bcc[4]) {
  const d3 n0 = cross3(t.M[1] - t.M[0], t.M[2] - t.M[0]);
  const d3 n1 = cross3(t.M[3] - t.M[0], t.M[1] - t.M[0]);
  const d3 n2 = cross3(t.M[3] - t.M[1], t.M[2] - t.M[1]);
  const d3 n3 = cross3(t.M[3] - t.M[2], t.M[0] - t.M[2]);
  const d3 p0 = p - t.M[0];
  double v[4];
  v[0] = dot3(p0, n0);
  v[1] = dot3(p0, n1);
  v[2] = dot3(p - t.M[1], n2);
  v[3] = dot3(p - t.M[2], n3);
  const double vol6 = dot3(t.M[3] - t.M[0], n0);
#pragma unroll
  for (int i = 0; i < 4; ++i) bcc[i] = -1;
  if (!(vol6 > 1.0e-20)) return false;
  const double inv = 1.0 / vol6;
#pragma unroll
  for (int i = 0; i < 4; ++i) bcc[i] = inv * v[i];
  return true;
}
// adjacency.hpp:136-158 barycentric_coords_tet (search_mesh_3d): vals scaled by 1/6, true volume
// (== measure_elements_real: both are tet_volume_from_basis(simplex_basis)); bcc stays 0 when
// vol < tol and the caller ignores the status
__device__ __forceinline__ void bcc_tet_s3d(const Tet& t, d3 p, double bcc[4], double tol) {
  const d3 n0 = cross3(t.M[1] - t.M[0], t.M[2] - t.M[0]);
  const d3 n1 = cross3(t.M[3] - t.M[0], t.M[1] - t.M[0]);
  const d3 n2 = cross3(t.M[3] - t.M[1], t.M[2] - t.M[1]);
  const d3 n3 = cross3(t.M[3] - t.M[2], t.M[0] - t.M[2]);
  const d3 p0 = p - t.M[0];
  double v[4];
  v[0] = 1.0 / 6.0 * dot3(p0, n0);
  v[1] = 1.0 / 6.0 * dot3(p0, n1);
  v[2] = 1.0 / 6.0 * dot3(p - t.M[1], n2);
  v[3] = 1.0 / 6.0 * dot3(p - t.M[2], n3);
#pragma unroll
  for (int i = 0; i < 4; ++i) bcc[i] = 0;
  if (t.vol < tol) return;
  const double inv = 1.0 / t.vol;
#pragma unroll
  for (int i = 0; i < 4; ++i) bcc[i] = inv * v[i];
}
// adjacency.tpp:23-39 barycentric_tri: edges (0,1),(1,2),(2,0)
__device__ __forceinline__ void bcc_tri(const Tri& t, d2 p, double bcc[3]) {
  bcc[0] = (cross2(t.M[1] - t.M[0], p - t.M[0]) / 2.0) / t.area;
  bcc[1] = (cross2(t.M[2] - t.M[1], p - t.M[1]) / 2.0) / t.area;
  bcc[2] = (cross2(t.M[0] - t.M[2], p - t.M[2]) / 2.0) / t.area;
}
template <int N>
__device__ __forceinline__ bool all_positive(const double* b, double tol) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < N; ++i) ok = ok && pp_gtez(b[i], tol);
  return ok;
}
// pumipic_utils.hpp:125-136 min_index (first strict minimum)
__device__ __forceinline__ int min_index4(const double* a) {
  int ind = 0;
  double mn = a[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (mn > a[i]) { mn = a[i]; ind = i; }
  return ind;
}
// pumipic_utils.hpp:138-149 max_index (first strict maximum)
__device__ __forceinline__ int max_index4(const double* a) {
  int ind = 0;
  double mx = a[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (mx < a[i]) { mx = a[i]; ind = i; }
  return ind;
}
// pumipic_utils.hpp:88-92 min3
__device__ __forceinline__ int min3(const double* a) {
  const bool first = a[0] < a[1];      // no dynamic indexing: a[] stays in registers
  const int idx = first ? 0 : 1;
  const double m = first ? a[0] : a[1];
  return (m < a[2]) ? idx : 2;
}

// ---------------------------------------------------------------- intersections
// Kokkos::min / Kokkos::max follow std::min / std::max (matters only for NaN operands)
__device__ __forceinline__ double kmin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double kmax(double a, double b) { return (a < b) ? b : a; }
// adjacency.tpp:152-178 ray_intersects_triangle; dir and seg_length hoisted by the caller
__device__ __forceinline__ bool ray_tri(d3 V0, d3 V1, d3 V2, d3 orig, d3 dir, double tol,
                                        int flip, d3& xp, double& dproj, double& closeness) {
  const d3 edge1 = (flip ? V1 : V2) - V0;   // faceVerts[2-flip]
  const d3 edge2 = (flip ? V2 : V1) - V0;   // faceVerts[flip+1]
  const d3 fnorm = cross3(edge2, edge1);
  const d3 pvec = cross3(dir, edge2);
  dproj = dot3(dir, fnorm);
  const double invdet = 1.0 / dproj;
  const d3 tvec = orig - V0;
  const double u = invdet * dot3(tvec, pvec);
  const d3 qvec = cross3(tvec, edge1);
  const double v = invdet * dot3(dir, qvec);
  const double t = invdet * dot3(edge2, qvec);
  xp = {orig.x + dir.x * t, orig.y + dir.y * t, orig.z + dir.z * t};
  closeness = kmax(kmax(kmin(fabs(u), fabs(1 - u)), kmin(fabs(v), fabs(1 - v))),
                   kmin(fabs(u + v), fabs(1 - u - v)));
  return (dproj >= tol) && (t >= -tol) && (u >= -tol) && (v >= -tol) && (u + v <= 1.0 + 2 * tol);
}
// adjacency.tpp:204-218 line_edge_2d
__device__ __forceinline__ bool line_edge(d2 E0, d2 E1, d2 orig, d2 dest, double tol, int flip,
                                          d2& xp) {
  const d2 A = flip ? E1 : E0;  // edgeVerts[vtx1 = flip]
  const d2 B = flip ? E0 : E1;  // edgeVerts[vtx2 = !flip]
  const d2 path = dest - orig;
  const d2 edge = B - A;
  const d2 nrm = {-edge.y, edge.x};
  const d2 nrmp = {-path.y, path.x};
  const double det = -dot2(nrm, path);
  const d2 rel = orig - A;
  const double s = dot2(nrmp, rel);
  const double t = dot2(nrm, rel);
  const double r = t / det;
  xp = {orig.x + r * path.x, orig.y + r * path.y};
  return det >= tol && s >= -tol && s <= det + tol && t >= -tol && t <= det + tol;
}
// adjacency.hpp:163-183 find_barycentric_tri_simple
__device__ __forceinline__ bool bcc_tri_simple(d3 a, d3 b, d3 c, d3 xp, double bc[3]) {
  const d3 ba = b - a, ca = c - a;
  d3 cr = cross3(ba, ca);
  cr = {cr.x * (1 / 2.0), cr.y * (1 / 2.0), cr.z * (1 / 2.0)};
  const double len = norm3(cr);
  const d3 nrm = {cr.x / len, cr.y / len, cr.z / len};
  const double area = dot3(nrm, cr);
  if (fabs(area) < 1e-20) return false;
  const double fac = 1 / (area * 2.0);
  const d3 xa = xp - a;
  bc[0] = fac * dot3(nrm, cross3(ba, xa));
  bc[1] = fac * dot3(nrm, cross3(c - b, xp - b));
  bc[2] = fac * dot3(nrm, cross3(xa, ca));
  return true;
}
// adjacency.hpp:230-273 line_triangle_intx_simple (dproj written only if both projections pass)
__device__ __forceinline__ bool line_tri_simple(d3 A, d3 B, d3 C, d3 origin, d3 dest, d3& xp,
                                                double& dproj, bool reverse, double tol) {
  xp = {0, 0, 0};
  bool found = false;
  const d3 line = dest - origin;
  d3 normv = cross3(B - A, C - A);
  if (reverse) normv = {-1 * normv.x, -1 * normv.y, -1 * normv.z};
  const double len = norm3(normv);
  const d3 unit = {normv.x / len, normv.y / len, normv.z / len};
  const double dist2plane = dot3(A - origin, unit);
  const double proj_end = dot3(unit, dest - A);
  if (dist2plane >= -tol && proj_end >= -tol) {
    dproj = dot3(line, unit);
    const double par_t = (dproj > 0) ? dist2plane / dproj : 0;
    xp = {origin.x + par_t * line.x, origin.y + par_t * line.y, origin.z + par_t * line.z};
    if (dproj > 0) {
      double bc[3];
      const bool res = bcc_tri_simple(A, B, C, xp, bc);
      if (res && bc[0] >= 0 && bc[0] <= 1 && bc[1] >= 0 && bc[1] <= 1 && bc[2] >= 0 && bc[2] <= 1)
        found = true;
    }
  }
  return found;
}

struct ThreadStats {
  int iters = 0, not_in = 0, not_found = 0, aborted = 0, active = 0, hops = 0;
};

// ---------------------------------------------------------------- the walk
template <int DIM, int MODE, bool PUSH>
__global__ void __launch_bounds__(128) k_search(SearchParams p) {
  using Rec = typename std::conditional<DIM == 3, Tet, Tri>::type;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  ThreadStats st;
  if (s < p.ps.capacity) {
    int erow;
    const bool mask = pp_slot_lookup(p.ps, s, erow);
    constexpr bool kNew = (MODE == M_BCC || MODE == M_RAY);
    int E = -1;
    bool live = false;
    if (mask) {
      if (kNew) {
        E = p.ids_empty ? erow : p.elem_ids[s];
        live = (E != -1);
      } else if (MODE == M_LEG2D) {          // adjacency.hpp:1045-1062
        // elem_ids_empty: the caller vouches that every id is -1 ("use the row element"); the
        // array is not read, so every structure kind and kernel variant starts identically
        E = p.ids_empty ? -1 : p.elem_ids[s];
        if (E == -1) E = erow;
        live = true;
        if (E == -p.nelems) { E = -1; live = false; }
      } else {                                // adjacency.hpp:586-598
        E = p.ids_empty ? erow : p.elem_ids[s];
        live = (E != -1);
      }
    }
    d3 xp_out = {0, 0, 0};
    int xface = -1;
    bool write_x = false;
    d3 tgt = {0, 0, 0}, org = {0, 0, 0};
    if (mask) {
      const bool from_orig = PUSH && p.push_from_orig;
      if ((live && MODE != M_LEG2D) || from_orig)
        org = {p.xo[s], p.xo[p.stride + s], p.xo[2 * p.stride + s]};
      // a masked particle is always pushed, even if it already left the domain (the reference's
      // push lambdas only test the mask)
      if (PUSH) {
        const d3 base = from_orig ? org : d3{p.xt[s], p.xt[p.stride + s], p.xt[2 * p.stride + s]};
        const d3 dr = {p.dir[s], p.dir[p.stride + s], p.dir[2 * p.stride + s]};
        tgt = {base.x + p.distance * dr.x, base.y + p.distance * dr.y, base.z + p.distance * dr.z};
        p.xt_rw[s] = tgt.x; p.xt_rw[p.stride + s] = tgt.y; p.xt_rw[2 * p.stride + s] = tgt.z;
      } else if (live) {
        tgt = {p.xt[s], p.xt[p.stride + s], p.xt[2 * p.stride + s]};
      }
    }
    if (live && kNew) {
      // finishUnmoved (adjacency.tpp:525-533): 3-component norm even in 2D
      if (norm3(tgt - org) < p.tol) live = false;
    }
    if (live) {
      st.active = 1;
      Rec rec;
      load_rec(p.walk, E, rec);
      bool done = false;
      if (kNew) {
        // check_initial_parents (adjacency.tpp:73-145)
        bool inside;
        if constexpr (DIM == 3) {
          double b[4];
          bcc_tet(rec, org, b);
          inside = all_positive<4>(b, p.tol);
        } else {
          double b[3];
          bcc_tri(rec, d2{org.x, org.y}, b);
          inside = all_positive<3>(b, p.tol);
        }
        if (!inside) { st.not_in = 1; E = -1; done = true; }
      }
      // hoisted ray direction (ray_intersects_triangle computes it per face from the same inputs)
      d3 dir = {0, 0, 0};
      if (MODE == M_RAY && DIM == 3) {
        const d3 disp = tgt - org;
        const double seg = norm3(disp);
        dir = {disp.x / seg, disp.y / seg, disp.z / seg};
      }
      int prevE = -1;
      int it = 0;
      while (!done) {
        ++it;
        int next = -1;
        if constexpr (MODE == M_BCC || MODE == M_LEG2D) {
          int f;
          int a;
          if constexpr (DIM == 3) {
            double b[4];
            bcc_tet(rec, tgt, b);
            done = all_positive<4>(b, kEps);
            f = min_index4(b);
            a = adj_of(rec, f);
          } else {
            double b[3];
            bcc_tri(rec, d2{tgt.x, tgt.y}, b);
            done = all_positive<3>(b, kEps);
            f = min3(b);
            a = adj_of(rec, f);
          }
          if (done) break;
          if (a < 0) { E = -1; done = true; break; }   // exposed side: particle leaves the domain
          next = a;
        } else if constexpr (MODE == M_RAY && DIM == 3) {
          // adjacency.tpp:316-361
          int exitf = -1, best = -1;
          double quality = -1;
#pragma unroll
          for (int fi = 0; fi < 4; ++fi) {
            const int a = rec.adj[fi];
            if (prevE >= 0 && a == prevE) continue;   // face_id == prevExit
            const unsigned code = (rec.codes >> (8 * fi)) & 0xffu;
            const d3 V0 = vert_of(rec, code & 3), V1 = vert_of(rec, (code >> 2) & 3),
                     V2 = vert_of(rec, (code >> 4) & 3);
            d3 xp;
            double dproj, closeness;
            const bool hit = ray_tri(V0, V1, V2, org, dir, p.tol, (code >> 6) & 1, xp, dproj, closeness);
            if (hit) { exitf = fi; xp_out = xp; write_x = true; }
            if (dproj > -p.tol && (quality < 0 || closeness < quality) && exitf == -1) {
              quality = closeness; best = fi; xp_out = xp; write_x = true;
            }
          }
          if (exitf == -1) exitf = best;
          if (exitf == -1) { done = true; break; }
          const int a = adj_of(rec, exitf);
          if (a < 0) { xface = -a - 1; done = true; break; }  // wall hit: element id kept
          next = a;
        } else if constexpr (MODE == M_RAY && DIM == 2) {
          // adjacency.tpp:285-313
          int exitf = -1;
#pragma unroll
          for (int ei = 0; ei < 3; ++ei) {
            const int a = rec.adj[ei];
            if (prevE >= 0 && a == prevE) continue;
            const unsigned code = (rec.codes >> (8 * ei)) & 0xffu;
            d2 xp;
            const bool hit = line_edge(vert_of(rec, code & 3), vert_of(rec, (code >> 2) & 3),
                                       d2{org.x, org.y}, d2{tgt.x, tgt.y}, p.tol, (code >> 6) & 1, xp);
            if (hit) { exitf = ei; xp_out = {xp.x, xp.y, 0}; write_x = true; }
          }
          if (exitf == -1) { done = true; break; }
          const int a = adj_of(rec, exitf);
          if (a < 0) { xface = -a - 1; done = true; break; }
          next = a;
        } else if constexpr (MODE == M_LEG3D) {
          // adjacency.hpp:607-740
          double b[4];
          if (it == 1) {
            bcc_tet_legacy(rec, org, b);
            if (!all_positive<4>(b, 1.0e-10)) st.aborted = 1;   // OMEGA_H_CHECK(false) :626
          }
          bcc_tet_legacy(rec, tgt, b);
          if (all_positive<4>(b, 1.0e-10)) { done = true; break; }
          double dproj[4] = {-1, -1, -1, -1};
          d3 xpts[4];
          bool intersected = false;
          bool decided = false;
#pragma unroll
          for (int fi = 0; fi < 4; ++fi) {
            if (decided) continue;
            const int a = rec.adj[fi];
            const unsigned code = (rec.codes >> (8 * fi)) & 0xffu;
            const d3 V0 = vert_of(rec, code & 3), V1 = vert_of(rec, (code >> 2) & 3),
                     V2 = vert_of(rec, (code >> 4) & 3);
            d3 xp;
            intersected = line_tri_simple(V0, V1, V2, org, tgt, xp, dproj[fi], (code >> 7) & 1, 1.0e-10);
            xpts[fi] = xp;
            if (intersected && a < 0) {
              done = true; xp_out = xp; write_x = true; xface = -a - 1; E = -1; decided = true;
            } else if (intersected) {
              next = a; decided = true;
            }
          }
          if (done) break;
          if (!intersected) {                       // :714-738
            const int mi = max_index4(dproj);
            const double dmax = mi == 0 ? dproj[0] : mi == 1 ? dproj[1] : mi == 2 ? dproj[2] : dproj[3];
            if (dmax >= 0) {
              const int a = adj_of(rec, mi);
              if (a < 0) {
                E = -1; xface = -a - 1; done = true; write_x = true;
                xp_out = mi == 0 ? xpts[0] : mi == 1 ? xpts[1] : mi == 2 ? xpts[2] : xpts[3];
                break;
              }
              const int fid = p.elem2sides[4 * (long)E + mi];   // reference bug reproduced (:726)
              if (fid < p.ndual) next = p.dual[fid];
              else { E = -1; done = true; break; }
            } else {
              E = -1; done = true; break;           // "leaked"
            }
          }
        } else if constexpr (MODE == M_S3D) {
          // adjacency.hpp:395-516: checkCurrentElm, findIntersection, processUndetected
          constexpr double tol3 = 1.0e-20;
          double b[4];
          if (it == 1) {                                   // checkParent :368-379 (row element)
            if (erow == E) {
              bcc_tet_s3d(rec, org, b, tol3);
            } else {
              Rec rrow;
              load_rec(p.walk, erow, rrow);
              bcc_tet_s3d(rrow, org, b, tol3);
            }
            if (!all_positive<4>(b, tol3)) st.aborted = 1;
          }
          bcc_tet_s3d(rec, tgt, b, tol3);
          if (all_positive<4>(b, tol3)) { done = true; break; }
          double dproj[4] = {-1, -1, -1, -1};
          d3 xpts[4];
          int ind_exp = -1, adj_f = -1;
#pragma unroll
          for (int fi = 0; fi < 4; ++fi) {
            const unsigned code = (rec.codes >> (8 * fi)) & 0xffu;
            const d3 V0 = vert_of(rec, code & 3), V1 = vert_of(rec, (code >> 2) & 3),
                     V2 = vert_of(rec, (code >> 4) & 3);
            const bool det = line_tri_simple(V0, V1, V2, org, tgt, xpts[fi], dproj[fi], (code >> 6) & 1, tol3);
            if (det && rec.adj[fi] < 0) ind_exp = fi;
            if (det && rec.adj[fi] >= 0) adj_f = fi;
          }
          if (ind_exp >= 0) {                              // wall collision :441-453
            xp_out = ind_exp == 0 ? xpts[0] : ind_exp == 1 ? xpts[1] : ind_exp == 2 ? xpts[2] : xpts[3];
            xface = -adj_of(rec, ind_exp) - 1; write_x = true;
            next = -1; done = true;
          }
          if (adj_f >= 0) {                                // interior; overrides a wall hit :456-468
            next = adj_of(rec, adj_f); done = false;
          }
          if (ind_exp < 0 && adj_f < 0) {                  // processUndetected :471-516
            const int mi = max_index4(dproj);
            const int a = adj_of(rec, mi);
            if (a < 0) {
              xp_out = mi == 0 ? xpts[0] : mi == 1 ? xpts[1] : mi == 2 ? xpts[2] : xpts[3];
              xface = -a - 1; write_x = true; next = -1; done = true;
            } else {
              const int fid = p.elem2sides[4 * (long)E + mi];   // dual indexed by face id (:510)
              if (fid < p.ndual) next = p.dual[fid];
              else { next = -1; done = true; }
            }
          }
          if (done) { E = next; break; }
        }
        // set_new_element + loop limit
        prevE = E;
        E = next;
        ++st.hops;
        if (MODE == M_LEG3D) {
          if (p.looplimit && it > p.looplimit) { st.not_found = 1; break; }     // :756
        } else if (MODE == M_S3D) {
          if (p.looplimit && it >= p.looplimit) { st.not_found = 1; break; }    // :528 (id kept)
        } else {
          if (p.looplimit && it >= p.looplimit) { st.not_found = 1; E = -1; break; }  // tpp:584-606
        }
        load_rec(p.walk, E, rec);
      }
      st.iters = it;
    }
    // ---- outputs
    if (mask || p.ids_empty || MODE == M_LEG2D || MODE == M_LEG3D || MODE == M_S3D) p.elem_ids[s] = mask ? E : -1;
    if (MODE == M_RAY) {
      // initializeIntersection resets every slot (tpp:535-549); hits overwrite
      p.inter_faces[s] = xface;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
        p.inter_points[(long)DIM * s + i] = write_x ? (i == 0 ? xp_out.x : i == 1 ? xp_out.y : xp_out.z) : 0.0;
    } else if ((MODE == M_LEG3D || MODE == M_S3D) && xface >= 0) {
      p.inter_faces[s] = xface;
      p.inter_points[3 * (long)s] = xp_out.x;
      p.inter_points[3 * (long)s + 1] = xp_out.y;
      p.inter_points[3 * (long)s + 2] = xp_out.z;
    }
  }
  // ---- warp-aggregated counters
  const unsigned full = 0xffffffffu;
  const int iters = __reduce_max_sync(full, st.iters);
  const int nin = __reduce_add_sync(full, st.not_in);
  const int nnf = __reduce_add_sync(full, st.not_found);
  const int nab = __reduce_add_sync(full, st.aborted);
  const int nac = __reduce_add_sync(full, st.active);
  const int nh = __reduce_add_sync(full, st.hops);
  if ((threadIdx.x & 31) == 0) {
    if (iters) atomicMax(&p.counters->max_iters, iters);
    if (nin) atomicAdd(&p.counters->not_in_elem, nin);
    if (nnf) atomicAdd(&p.counters->not_found, nnf);
    if (nab) atomicAdd(&p.counters->aborted, nab);
    if (nac) atomicAdd(&p.counters->active, nac);
    if (nh) atomicAdd(&p.counters->hops, (unsigned long long)nh);
  }
}

// ---------------------------------------------------------------- staged BCC walk
// The barycentric walk (search_mesh with requireIntersection=false, and search_mesh_2d) as a
// block-synchronous pipeline.  A block owns BLOCK consecutive slots.  Every round:
//   1. the block fetches the walk records of all queued particles COOPERATIVELY: 16-byte piece
//      k of record i is loaded by thread i*PIECES+k, so each record costs one or two coalesced
//      L1 wavefronts instead of one wavefront per piece per lane;
//   2. records are staged in shared memory (stride chosen bank-conflict free for 128-bit reads);
//   3. each queued particle is evaluated by one thread; particles that must hop are pushed
//      into a shared-memory queue with a warp-aggregated slot claim, so the next round runs on
//      densely packed warps (walk lengths differ: ~49 % stop in round 0, ~2 % need >= 4 hops).
template <int DIM> struct StageCfg;
template <> struct StageCfg<3> {
  static constexpr int PIECES = 12, STRIDE = 208;   // 52 words: 8 lanes x 4 banks, conflict free
  using Raw = PPBccRec3;
};
template <> struct StageCfg<2> {
  static constexpr int PIECES = 5, STRIDE = 80;     // 20 words: conflict free as well
  using Raw = PPTriRec;
};

struct Bcc3 {
  d3 a0, a1, a2, n0, n1, n2, n3;
  double inv_vol;
  int adj[4];
};
__device__ __forceinline__ void read_stage(const unsigned char* st, Bcc3& r) {
  const double2* q = reinterpret_cast<const double2*>(st);
  const double2 p0 = q[0], p1 = q[1], p2 = q[2], p3 = q[3], p4 = q[4], p5 = q[5];
  const double2 p6 = q[6], p7 = q[7], p8 = q[8], p9 = q[9], p10 = q[10];
  const int4 p11 = *reinterpret_cast<const int4*>(q + 11);
  r.a0 = {p0.x, p0.y, p1.x};
  r.a1 = {p1.y, p2.x, p2.y};
  r.a2 = {p3.x, p3.y, p4.x};
  r.n0 = {p4.y, p5.x, p5.y};
  r.n1 = {p6.x, p6.y, p7.x};
  r.n2 = {p7.y, p8.x, p8.y};
  r.n3 = {p9.x, p9.y, p10.x};
  r.inv_vol = p10.y;
  r.adj[0] = p11.x; r.adj[1] = p11.y; r.adj[2] = p11.z; r.adj[3] = p11.w;
}
__device__ __forceinline__ void read_stage(const unsigned char* st, Tri& t) {
  const double2* q = reinterpret_cast<const double2*>(st);
  const double2 a0 = q[0], a1 = q[1], a2 = q[2];
  const int4 q3 = *reinterpret_cast<const int4*>(q + 3);
  const int4 q4 = *reinterpret_cast<const int4*>(q + 4);
  t.M[0] = {a0.x, a0.y};
  t.M[1] = {a1.x, a1.y};
  t.M[2] = {a2.x, a2.y};
  t.area = __hiloint2double(q3.y, q3.x);
  t.adj[0] = q3.z; t.adj[1] = q3.w; t.adj[2] = q4.x;
  t.codes = (unsigned)q4.y;
  t.cls = q4.z;
  t.aux = q4.w;
}
// barycentric_tet with the particle-independent half read from the record (bit-identical)
__device__ __forceinline__ void bcc_tet(const Bcc3& t, d3 p, double bcc[4]) {
  const d3 p0 = p - t.a0;
  double v[4];
  v[0] = dot3(p0, t.n0);
  v[1] = dot3(p0, t.n1);
  v[2] = dot3(p - t.a1, t.n2);
  v[3] = dot3(p - t.a2, t.n3);
#pragma unroll
  for (int i = 0; i < 4; ++i) bcc[i] = t.inv_vol * v[i];
  if (__builtin_expect(!(t.inv_vol > 0), 0)) {    // barycentric_tet fails on a degenerate tet
#pragma unroll
    for (int i = 0; i < 4; ++i) bcc[i] = -1;
  }
}

template <int DIM, int BLOCK>
__device__ __forceinline__ void stage_fetch(const typename StageCfg<DIM>::Raw* table,
                                            const int* q_E, int n, unsigned char* stage) {
  using Cfg = StageCfg<DIM>;
  constexpr int HALF = (Cfg::PIECES + 1) / 2;
  const int total = n * Cfg::PIECES;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int4 v[HALF];
    int dst[HALF];
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
      const int idx = threadIdx.x + (h * HALF + k) * BLOCK;
      dst[k] = -1;
      if (h * HALF + k < Cfg::PIECES && idx < total) {
        const int item = idx / Cfg::PIECES;
        const int piece = idx - item * Cfg::PIECES;
        const int e = q_E[item];
        if (e >= 0) {
          v[k] = __ldg(reinterpret_cast<const int4*>(table + e) + piece);
          dst[k] = item * Cfg::STRIDE + piece * 16;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < HALF; ++k)
      if (dst[k] >= 0) *reinterpret_cast<int4*>(stage + dst[k]) = v[k];
  }
}

template <int DIM, bool LEG2D, bool PUSH, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (DIM == 3 ? 3 : 4)) k_walk_bcc(SearchParams p) {
  using Cfg = StageCfg<DIM>;
  using Rec = typename std::conditional<DIM == 3, Bcc3, Tri>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* stage = smem;
  double* q_tx = reinterpret_cast<double*>(smem + BLOCK * Cfg::STRIDE);
  double* q_ty = q_tx + BLOCK;
  double* q_tz = q_ty + BLOCK;
  int* q_E = reinterpret_cast<int*>(q_tz + BLOCK);
  int* q_slot = q_E + BLOCK;
  int* q_it = q_slot + BLOCK;
  __shared__ int q_count;
  const auto* table = reinterpret_cast<const typename Cfg::Raw*>(DIM == 3 ? p.walk_bcc : p.walk);

  const int tid = threadIdx.x;
  const int slot = blockIdx.x * BLOCK + tid;
  if (tid == 0) q_count = 0;
  ThreadStats st;
  int E = -1;
  bool live = false;
  d3 tgt = {0, 0, 0}, org = {0, 0, 0};
  // ---- setup: setInitial / fill, push, finishUnmoved (same rules as k_search)
  if (slot < p.ps.capacity) {
    int erow;
    const bool mask = pp_slot_lookup(p.ps, slot, erow);
    if (mask) {
      if (!LEG2D) {
        E = p.ids_empty ? erow : p.elem_ids[slot];
        live = (E != -1);
      } else {                                  // adjacency.hpp:1045-1062
        E = p.ids_empty ? -1 : p.elem_ids[slot];   // see k_search
        if (E == -1) E = erow;
        live = true;
        if (E == -p.nelems) { E = -1; live = false; }
      }
      const bool from_orig = PUSH && p.push_from_orig;
      if ((live && !LEG2D) || from_orig)
        org = {p.xo[slot], p.xo[p.stride + slot], p.xo[2 * p.stride + slot]};
      if (PUSH) {
        const d3 base = from_orig ? org
                                  : d3{p.xt[slot], p.xt[p.stride + slot], p.xt[2 * p.stride + slot]};
        const d3 dr = {p.dir[slot], p.dir[p.stride + slot], p.dir[2 * p.stride + slot]};
        tgt = {base.x + p.distance * dr.x, base.y + p.distance * dr.y, base.z + p.distance * dr.z};
        p.xt_rw[slot] = tgt.x; p.xt_rw[p.stride + slot] = tgt.y; p.xt_rw[2 * p.stride + slot] = tgt.z;
      } else if (live) {
        tgt = {p.xt[slot], p.xt[p.stride + slot], p.xt[2 * p.stride + slot]};
      }
      if (live && !LEG2D && norm3(tgt - org) < p.tol) live = false;   // finishUnmoved
    }
    if (!live && (mask || p.ids_empty || LEG2D)) p.elem_ids[slot] = mask ? E : -1;
  }
  q_E[tid] = live ? E : -1;
  __syncthreads();
  // ---- round 0: every live slot is an item; origin check + first target test share a record
  stage_fetch<DIM, BLOCK>(table, q_E, BLOCK, stage);
  __syncthreads();
  auto advance = [&](const Rec& rec, int myslot, int& e, int it, d3 t) {
    // one walk iteration: find_exit_face (BCC) + check_model_intersection + set_new_element
    bool done;
    int f;
    if constexpr (DIM == 3) {
      double b[4];
      bcc_tet(rec, t, b);
      done = all_positive<4>(b, kEps);
      f = min_index4(b);
    } else {
      double b[3];
      bcc_tri(rec, d2{t.x, t.y}, b);
      done = all_positive<3>(b, kEps);
      f = min3(b);
    }
    bool push = false;
    int next = -1;
    if (!done) {
      const int a = adj_of(rec, f);
      if (a < 0) {
        e = -1;                       // exposed side: the particle leaves the domain
      } else {
        ++st.hops;
        if (p.looplimit && it >= p.looplimit) { st.not_found = 1; e = -1; }  // tpp:584-606
        else { push = true; next = a; }
      }
    }
    // warp-aggregated claim of queue positions
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, push);
    if (m) {
      const int leader = __ffs(m) - 1;
      int base = 0;
      if ((tid & 31) == leader) base = atomicAdd(&q_count, __popc(m));
      base = __shfl_sync(act, base, leader);
      if (push) {
        const int pos = base + __popc(m & ((1u << (tid & 31)) - 1u));
        q_E[pos] = next; q_slot[pos] = myslot; q_it[pos] = it;
        q_tx[pos] = t.x; q_ty[pos] = t.y; q_tz[pos] = t.z;
      }
    }
    if (!push) p.elem_ids[myslot] = e;
    st.iters = it > st.iters ? it : st.iters;
  };
  if (live) {
    st.active = 1;
    Rec rec;
    read_stage(stage + tid * Cfg::STRIDE, rec);
    bool inside = true;
    if (!LEG2D) {                                 // check_initial_parents (tpp:73-145)
      if constexpr (DIM == 3) {
        double b[4];
        bcc_tet(rec, org, b);
        inside = all_positive<4>(b, p.tol);
      } else {
        double b[3];
        bcc_tri(rec, d2{org.x, org.y}, b);
        inside = all_positive<3>(b, p.tol);
      }
    }
    if (!inside) {
      st.not_in = 1;
      p.elem_ids[slot] = -1;
    } else {
      advance(rec, slot, E, 1, tgt);
    }
  }
  // ---- rounds >= 1 on the compacted queue
  while (true) {
    __syncthreads();
    const int n = q_count;
    if (n == 0) break;
    const bool has = tid < n;
    int myslot = 0, it = 0;
    if (has) {
      E = q_E[tid]; myslot = q_slot[tid]; it = q_it[tid];
      tgt = {q_tx[tid], q_ty[tid], q_tz[tid]};
    }
    __syncthreads();
    if (tid == 0) q_count = 0;
    stage_fetch<DIM, BLOCK>(table, q_E, n, stage);
    __syncthreads();
    if (has) {
      Rec rec;
      read_stage(stage + tid * Cfg::STRIDE, rec);
      advance(rec, myslot, E, it + 1, tgt);
    }
  }
  // ---- warp-aggregated counters
  const unsigned full = 0xffffffffu;
  const int iters = __reduce_max_sync(full, st.iters);
  const int nin = __reduce_add_sync(full, st.not_in);
  const int nnf = __reduce_add_sync(full, st.not_found);
  const int nac = __reduce_add_sync(full, st.active);
  const int nh = __reduce_add_sync(full, st.hops);
  if ((tid & 31) == 0) {
    if (iters) atomicMax(&p.counters->max_iters, iters);
    if (nin) atomicAdd(&p.counters->not_in_elem, nin);
    if (nnf) atomicAdd(&p.counters->not_found, nnf);
    if (nac) atomicAdd(&p.counters->active, nac);
    if (nh) atomicAdd(&p.counters->hops, (unsigned long long)nh);
  }
}

// ---------------------------------------------------------------- chunk walk (Sell-C-sigma, C = 32)
// The barycentric search_mesh for a Sell-C-sigma structure whose elem_ids are seeded from the
// rows (every reference call site: the particles were rebuilt into the row of their element).
//
//   * warp per chunk: lane r owns row r of the chunk, so the walk record of the row's element is
//     gathered ONCE per chunk into registers and reused for every column of the row; the
//     particle columns themselves (x, dir, xtgt, elem_ids) are fully coalesced 256 B accesses;
//   * origin check (check_initial_parents) and the first exit test use that register record;
//   * particles that must hop are pushed, with a warp-aggregated claim, into a per-warp
//     shared-memory queue; whenever the queue holds a full warp of work it is drained: the 32
//     neighbour records are fetched COOPERATIVELY with 16-byte cp.async pieces (consecutive lanes
//     read consecutive pieces, so a record costs 1.5 L1 wavefronts instead of 12) into a
//     bank-conflict-free stage, each lane evaluates one queued particle and re-queues it if it
//     must hop again.  Lanes therefore stay full although walk lengths differ (49 % stop at
//     once, 2 % need four hops or more);
//   * chunks are handed out by an atomic counter, there is no block-level synchronisation.
#ifndef PP_SCS_MINB
#define PP_SCS_MINB 5        // resident 128-thread blocks per SM the chunk walk is compiled for (20 warps: measured +6.6 % over 4)
#endif
#ifndef PP_SCS_RING
#define PP_SCS_RING 1        // depth 1..3 measured equal at 16 warps/SM; 1 leaves shared memory for the fifth block
#endif
constexpr int kQCap = 64;    // queue entries per warp; a full warp of work is drained at once
constexpr int kRing = PP_SCS_RING;   // particle columns in flight per warp (cp.async ring)

template <int DIM>
struct WarpSmem {
  static constexpr int STAGE = 32 * StageCfg<DIM>::STRIDE;
  static constexpr int QUEUE = kQCap * (3 * 8 + 3 * 4);
  static constexpr int RING = kRing * 6 * 32 * 8;
  static constexpr int BYTES = STAGE + QUEUE + RING + 32 * 4 * 4;   // + row adjacency (4 ints per lane)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int DIM>
__device__ __forceinline__ void warp_stage_fetch(const typename StageCfg<DIM>::Raw* table,
                                                 const int* q_E, int n, unsigned char* stage,
                                                 int lane) {
  using Cfg = StageCfg<DIM>;
  if constexpr (Cfg::PIECES == 12) {
    // piece index lane + 32k: 32k mod 12 has period 3 in k, so (item, piece) of k = 3m + r is
    // (item_r + 8m, piece_r): three divisions per call instead of twelve
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int idx = lane + 32 * r;
      const int item_r = idx / 12;
      const int piece_r = idx - item_r * 12;
      const unsigned char* src_r = reinterpret_cast<const unsigned char*>(table) + piece_r * 16;
      unsigned char* dst_r = stage + item_r * Cfg::STRIDE + piece_r * 16;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int item = item_r + 8 * m;
        if (item < n) cp_async16(dst_r + m * 8 * Cfg::STRIDE, src_r + (long)q_E[item] * (long)sizeof(typename Cfg::Raw));
      }
    }
  } else {
    const int total = n * Cfg::PIECES;
#pragma unroll
    for (int k = 0; k < Cfg::PIECES; ++k) {
      const int idx = lane + k * 32;
      if (idx < total) {
        const int item = idx / Cfg::PIECES;
        const int piece = idx - item * Cfg::PIECES;
        const int e = q_E[item];
        cp_async16(stage + item * Cfg::STRIDE + piece * 16, reinterpret_cast<const int4*>(table + e) + piece);
      }
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
}

// PUSH: 0 no push, 1 xt += d*dir (test_adj form), 2 xt = xo + d*dir (PIC form; the bench's path)
template <int DIM, int PUSH, bool LEG, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, PP_SCS_MINB) k_walk_scs(SearchParams p) {
  using Cfg = StageCfg<DIM>;
  using Rec = typename std::conditional<DIM == 3, Bcc3, Tri>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  unsigned char* wbase = smem + (threadIdx.x >> 5) * WarpSmem<DIM>::BYTES;
  unsigned char* stage = wbase;
  double* q_tx = reinterpret_cast<double*>(wbase + WarpSmem<DIM>::STAGE);
  double* q_ty = q_tx + kQCap;
  double* q_tz = q_ty + kQCap;
  int* q_E = reinterpret_cast<int*>(q_tz + kQCap);
  int* q_slot = q_E + kQCap;
  int* q_it = q_slot + kQCap;
  double* ring = reinterpret_cast<double*>(wbase + WarpSmem<DIM>::STAGE + WarpSmem<DIM>::QUEUE) + lane;
  int* q_rowE = reinterpret_cast<int*>(wbase + WarpSmem<DIM>::STAGE + WarpSmem<DIM>::QUEUE + WarpSmem<DIM>::RING);
  const auto* table = reinterpret_cast<const typename Cfg::Raw*>(DIM == 3 ? p.walk_bcc : p.walk);
  const unsigned full = 0xffffffffu;
  const unsigned lt = (1u << lane) - 1u;
  constexpr int kAdjOff = DIM == 3 ? 176 : 56;   // byte offset of adj[] inside a staged record
  int* row_adj = q_rowE + lane * 4;              // adjacency of this lane's row (aliases q_rowE)
  int st_iters = 0, st_active = 0;               // per-lane counters; the rest are warp-uniform
  int n_push = 0, n_lost = 0, n_notin = 0;
  int qn = 0;   // queue fill, warp-uniform

  // one walk iteration of a particle whose record is `rec` (adjacency read from shared memory at
  // `adj`): find_exit_face (BCC) + check_model_intersection + set_new_element.  Returns true if
  // the particle must hop to `next`.
  auto advance = [&](const Rec& rec, const int* adj, int& e, int it, d3 t, int& next, bool& lost) -> bool {
    bool done;
    int f;
    if constexpr (DIM == 3) {
      double b[4];
      bcc_tet(rec, t, b);
      done = all_positive<4>(b, kEps);
      f = min_index4(b);
    } else {
      double b[3];
      bcc_tri(rec, d2{t.x, t.y}, b);
      done = all_positive<3>(b, kEps);
      f = min3(b);
    }
    st_iters = it > st_iters ? it : st_iters;
    if (done) return false;
    const int a = adj[f];
    if (a < 0) { e = -1; return false; }          // exposed side: the particle leaves the domain
    if (p.looplimit && it >= p.looplimit) { lost = true; e = -1; return false; }  // tpp:584-606
    next = a;
    return true;
  };
  // hops = particles that moved to a neighbour (queued) + those stopped there by the loop limit
  auto enqueue = [&](bool push, bool lost, int slot, int next, int it, d3 t) {
    const unsigned m = __ballot_sync(full, push);
    if (p.looplimit) n_lost += __popc(__ballot_sync(full, lost));
    n_push += __popc(m);
    if (push) {
      const int pos = qn + __popc(m & lt);
      q_E[pos] = next; q_slot[pos] = slot; q_it[pos] = it;
      q_tx[pos] = t.x; q_ty[pos] = t.y; q_tz[pos] = t.z;
    }
    qn += __popc(m);
    __syncwarp();
  };
  auto drain = [&](int n) {       // n <= 32 entries from the top of the queue
    const int base = qn - n;
    const bool has = lane < n;
    int E = -1, slot = 0, it = 0;
    d3 t = {0, 0, 0};
    if (has) {
      E = q_E[base + lane]; slot = q_slot[base + lane]; it = q_it[base + lane];
      t = {q_tx[base + lane], q_ty[base + lane], q_tz[base + lane]};
    }
    warp_stage_fetch<DIM>(table, q_E + base, n, stage, lane);
    qn = base;
    bool push = false, lost = false;
    int next = -1;
    if (has) {
      Rec rec;
      read_stage(stage + lane * Cfg::STRIDE, rec);
      push = advance(rec, reinterpret_cast<const int*>(stage + lane * Cfg::STRIDE + kAdjOff), E, it + 1,
                     t, next, lost);
      if (!push) p.elem_ids[slot] = E;
    }
    __syncwarp();                 // every lane has read its stage row / queue entry
    enqueue(push, lost, slot, next, it + 1, t);
  };

  // row records of a chunk: cooperative fetch through the stage, then one register copy per lane
  auto fetch_rows = [&](int e, Rec& rec) {
    q_rowE[lane] = e;
    __syncwarp();
    warp_stage_fetch<DIM>(table, q_rowE, 32, stage, lane);
    read_stage(stage + lane * Cfg::STRIDE, rec);
    const int* sa = reinterpret_cast<const int*>(stage + lane * Cfg::STRIDE + kAdjOff);
    row_adj[0] = sa[0]; row_adj[1] = sa[1]; row_adj[2] = sa[2];
    if (DIM == 3) row_adj[3] = sa[3];
    __syncwarp();
  };
  constexpr bool from_orig = PUSH == 2;
  const double* __restrict__ colA = p.xo;                 // origin
  const double* __restrict__ colB = PUSH ? p.dir : p.xt;  // direction | target
  // column prefetch: each lane copies its own six doubles into its private ring entries, so the
  // only synchronisation is its own cp.async.wait_group
  auto issue_col = [&](int s, bool m, int r) {
    if (m) {
      double* d = ring + r * (6 * 32);
#pragma unroll
      for (int k = 0; k < (LEG ? DIM : 3); ++k) {   // search_mesh_2d reads the target only
        if (!LEG) cp_async8(d + k * 32, colA + k * p.stride + s);
        cp_async8(d + (3 + k) * 32, colB + k * p.stride + s);
      }
    }
    cp_async_commit();
  };

  // Work unit = one vertical slice of a chunk (at most V columns, SellCSigma.h:26-60), not a whole
  // chunk: a row that holds a large share of the particles (pseudoXGCm's load puts the rounding
  // shortfall of ~1 M particles into ONE element, test/pseudoXGCm.cpp:253-263) is spread over the
  // warps instead of being walked by one.  Slices of chunks [chunk_begin, chunk_end): the slices
  // are sorted by chunk, the range is found by bisection once per warp.
  // When every chunk fits one slice (the usual case) the unit is the chunk itself.
  const int* __restrict__ s2c = p.ps.slice_to_chunk;
  const int* __restrict__ soff = p.ps.offsets;
  const int* __restrict__ cstart = p.ps.chunk_start;
  const bool by_slice = p.ps.sliced != 0;
  int unit_lo = p.chunk_begin, unit_hi = p.chunk_end;
  if (by_slice) {
    unit_lo = 0; unit_hi = p.ps.nslices;
    if (p.chunk_begin > 0 || p.chunk_end < p.ps.nchunks) {
      int lo = 0, hi = p.ps.nslices;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(s2c + mid) < p.chunk_begin) lo = mid + 1; else hi = mid; }
      unit_lo = lo;
      hi = p.ps.nslices;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(s2c + mid) < p.chunk_end) lo = mid + 1; else hi = mid; }
      unit_hi = lo;
    }
  }
  while (true) {
    int sl = 0;
    if (lane == 0) sl = unit_lo + atomicAdd(p.sched, 1);
    sl = __shfl_sync(full, sl, 0);
    if (sl >= unit_hi) break;
    const int c = by_slice ? __ldg(s2c + sl) : sl;
    const int s0 = by_slice ? __ldg(soff + sl) : __ldg(cstart + c);
    const int s1 = by_slice ? __ldg(soff + sl + 1) : __ldg(cstart + c + 1);
    if (s1 <= s0) continue;
    const int ncols = (s1 - s0) >> 5;
    const int rowE = __ldg(p.ps.row_to_element + c * 32 + lane);
    const int rowF = rowE < p.nelems ? rowE : 0;  // padding rows of the last chunk hold no particle
    for (int cb = 0; cb < ncols; cb += 32) {      // batches of 32 columns: one mask word per lane
      const int nb = ncols - cb < 32 ? ncols - cb : 32;
      const uint32_t mw = lane < nb ? __ldg(p.ps.mask_bits + (s0 >> 5) + cb + lane) : 0u;
      const unsigned nz = __ballot_sync(full, mw != 0u);
      const int nlive = 32 - __clz(nz);           // rows fill from column 0: empty columns trail
      // my_cols: bit j set iff this lane's slot in column j holds a particle
      uint32_t my_cols = 0;
      for (int j = 0; j < nlive; ++j) my_cols |= ((__shfl_sync(full, mw, j) >> lane) & 1u) << j;
      const int sbase = s0 + cb * 32 + lane;
#pragma unroll
      for (int k = 0; k < kRing; ++k) issue_col(sbase + k * 32, (my_cols >> k) & 1u, k);
      for (int j = nlive; j < nb; ++j) p.elem_ids[sbase + j * 32] = -1;
      Rec rec;
      fetch_rows(rowF, rec);                      // also waits for the first columns
      int r = 0;
      for (int j = 0; j < nlive; ++j) {
        const int s = sbase + j * 32;
        cp_async_wait<kRing - 1>();
        const bool mask = (my_cols >> j) & 1u;
        const double* d = ring + r * (6 * 32);
        d3 org = {0, 0, 0}, aux = {0, 0, 0};
        if (mask) {
          if (!LEG) org = {d[0], d[32], d[64]};
          aux = {d[96], d[128], (LEG && DIM == 2) ? 0.0 : d[160]};
        }
        issue_col(s + kRing * 32, j + kRing < nlive && ((my_cols >> ((j + kRing) & 31)) & 1u), r);
        r = r + 1 == kRing ? 0 : r + 1;
        int E = -1;
        bool push = false, lost = false, notin = false;
        int next = -1;
        d3 tgt = {0, 0, 0};
        if (mask) {
          E = rowE;                               // setInitial (tpp:504-515)
          if (PUSH) {
            const d3 base = from_orig ? org : d3{p.xt[s], p.xt[p.stride + s], p.xt[2 * p.stride + s]};
            tgt = {base.x + p.distance * aux.x, base.y + p.distance * aux.y, base.z + p.distance * aux.z};
            p.xt_rw[s] = tgt.x; p.xt_rw[p.stride + s] = tgt.y; p.xt_rw[2 * p.stride + s] = tgt.z;
          } else {
            tgt = aux;
          }
          const d3 mv = tgt - org;
          if (LEG) {                              // search_mesh_2d (adjacency.hpp:1045-1117): no origin test
            st_active += 1;
            push = advance(rec, row_adj, E, 1, tgt, next, lost);
          } else if (!(dot3(mv, mv) < p.unmoved_sq)) {   // finishUnmoved (tpp:525-533), see unmoved_threshold()
            st_active += 1;
            bool inside;                          // check_initial_parents (tpp:73-145)
            if constexpr (DIM == 3) {
              double b[4];
              bcc_tet(rec, org, b);
              inside = all_positive<4>(b, p.tol);
            } else {
              double b[3];
              bcc_tri(rec, d2{org.x, org.y}, b);
              inside = all_positive<3>(b, p.tol);
            }
            if (!inside) { notin = true; E = -1; }
            else push = advance(rec, row_adj, E, 1, tgt, next, lost);
          }
        }
        if (!push) p.elem_ids[s] = E;             // unmasked slots get -1 (elem_ids is seeded here)
        if (__any_sync(full, notin)) n_notin += __popc(__ballot_sync(full, notin));
        enqueue(push, lost, s, next, 1, tgt);
        while (qn >= 32) drain(32);
      }
    }
  }
  while (qn > 0) drain(qn < 32 ? qn : 32);

  // ---- warp-aggregated counters
  const int iters = __reduce_max_sync(full, st_iters);
  const int nac = __reduce_add_sync(full, st_active);
  if (lane == 0) {
    if (iters) atomicMax(&p.counters->max_iters, iters);
    if (n_notin) atomicAdd(&p.counters->not_in_elem, n_notin);
    if (n_lost) atomicAdd(&p.counters->not_found, n_lost);
    if (nac) atomicAdd(&p.counters->active, nac);
    if (n_push + n_lost) atomicAdd(&p.counters->hops, (unsigned long long)(n_push + n_lost));
  }
}

// finishUnmoved tests sqrt(d) < tol with d = |tgt - org|^2.  sqrt is correctly rounded and
// monotone, so the test equals d < T with T the smallest double whose square root is >= tol;
// T is found here on the host and the kernel skips the square root (bit-identical decisions).
double unmoved_threshold(double tol) {
  if (!(tol > 0)) return 0.0;             // sqrt(d) < tol is never true
  double T = tol * tol;
  while (sqrt(T) >= tol && T > 0) T = nextafter(T, 0.0);
  while (sqrt(T) < tol) T = nextafter(T, INFINITY);
  return T;
}

int g_sm_count = 0;

// Optional L2 access-policy window over the walk table (pp_search_set_l2_window / PUMIPIC_L2_WINDOW):
// a fraction of the table's lines is marked persisting, everything else the kernel touches (the
// particle columns, read once) streams.  Off by default: the fused kernel's DRAM traffic is already
// within 2 % of its floor (every record comes from DRAM about once, profiles/r2l_ncu_k_walk_scs.txt),
// so there is nothing for the window to save, and the carve-out shrinks the L2 everything else uses:
// measured 0.252 -> 0.464 ms per 10 M particles (profiles/r2F_l2_window_ab.txt).
double g_l2_window = -1.0;      // < 0: read the environment on first use
size_t g_l2_persist_max = 0, g_l2_window_max = 0;

template <class K>
pp_status launch_with_window(K kernel, int grid, int block, size_t smem, cudaStream_t s, const SearchParams& p,
                             const void* table, size_t table_bytes) {
  if (g_l2_window < 0) {
    const char* env = getenv("PUMIPIC_L2_WINDOW");
    g_l2_window = env ? atof(env) : 0.0;
    if (g_l2_window < 0) g_l2_window = 0;
  }
  if (g_l2_window <= 0 || !table || !table_bytes) {
    kernel<<<grid, block, smem, s>>>(p);
    return PP_OK;
  }
  if (!g_l2_persist_max) {
    int dev = 0, v = 0;
    PP_CUDA(cudaGetDevice(&dev));
    PP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev));
    g_l2_persist_max = (size_t)v;
    PP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    g_l2_window_max = (size_t)v;
    if (g_l2_persist_max) PP_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_l2_persist_max));
  }
  if (!g_l2_persist_max || !g_l2_window_max) {      // no persisting L2 on this device
    kernel<<<grid, block, smem, s>>>(p);
    return PP_OK;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
  cudaAccessPolicyWindow& w = attr[0].val.accessPolicyWindow;
  w.base_ptr = const_cast<void*>(table);
  w.num_bytes = table_bytes < g_l2_window_max ? table_bytes : g_l2_window_max;
  // the share of the window's lines that may persist: the requested fraction of the persisting carve-out
  const double ratio = g_l2_window * (double)g_l2_persist_max / (double)w.num_bytes;
  w.hitRatio = (float)(ratio < 1.0 ? ratio : 1.0);
  w.hitProp = cudaAccessPropertyPersisting;
  w.missProp = cudaAccessPropertyStreaming;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem; cfg.stream = s; cfg.attrs = attr; cfg.numAttrs = 1;
  PP_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  return PP_OK;
}

template <int DIM, bool LEG>
pp_status launch_walk_scs(const SearchParams& p, bool push, cudaStream_t s) {
  constexpr int WARPS = 4;
  constexpr size_t smem = (size_t)WARPS * WarpSmem<DIM>::BYTES;
  if (!g_sm_count) {
    int dev = 0;
    PP_CUDA(cudaGetDevice(&dev));
    PP_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int want = pp_div_up(p.ps.sliced ? p.ps.nslices : p.chunk_end - p.chunk_begin, WARPS);
  if (want <= 0) return PP_OK;                         // nothing but empty chunks in the range
  const int persistent = g_sm_count * PP_SCS_MINB;
  const int grid = want < persistent ? want : persistent;
  const void* table = DIM == 3 ? p.walk_bcc : p.walk;
  const size_t table_bytes = (size_t)p.nelems * sizeof(typename StageCfg<DIM>::Raw);
  if (push && !LEG && p.push_from_orig) {
    auto k = k_walk_scs<DIM, 2, false, WARPS>;
    PP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_TRY(launch_with_window(k, grid, WARPS * 32, smem, s, p, table, table_bytes));
  } else if (push && !LEG) {
    auto k = k_walk_scs<DIM, 1, false, WARPS>;
    PP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_TRY(launch_with_window(k, grid, WARPS * 32, smem, s, p, table, table_bytes));
  } else {
    auto k = k_walk_scs<DIM, 0, LEG, WARPS>;
    PP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_TRY(launch_with_window(k, grid, WARPS * 32, smem, s, p, table, table_bytes));
  }
  return PP_OK;
}

template <int DIM, bool LEG2D>
pp_status launch_walk_bcc(const SearchParams& p, bool push, cudaStream_t s) {
  constexpr int BLOCK = 256;
  constexpr size_t smem = (size_t)BLOCK * StageCfg<DIM>::STRIDE + (size_t)BLOCK * (3 * 8 + 3 * 4);
  const int grid = pp_div_up(p.ps.capacity, BLOCK);
  if (push) {
    auto k = k_walk_bcc<DIM, LEG2D, true, BLOCK>;
    PP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, BLOCK, smem, s>>>(p);
  } else {
    auto k = k_walk_bcc<DIM, LEG2D, false, BLOCK>;
    PP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, BLOCK, smem, s>>>(p);
  }
  return PP_OK;
}

template <int DIM, int MODE>
void launch(const SearchParams& p, bool push, cudaStream_t s) {
  const int block = 128;
  const int grid = pp_div_up(p.ps.capacity, block);
  if (push) k_search<DIM, MODE, true><<<grid, block, 0, s>>>(p);
  else k_search<DIM, MODE, false><<<grid, block, 0, s>>>(p);
}

int g_staged_walk = 2;

pp_status read_stats(pp_mesh* mesh, int variant, int looplimit, pp_search_stats* out, cudaStream_t s) {
  SearchCounters h;
  PP_CUDA(cudaMemcpyAsync(&h, mesh->stats_dev, sizeof(h), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  out->not_in_elem = h.not_in_elem;
  out->not_found = h.not_found;
  out->aborted = h.aborted;
  out->active = h.active;
  out->hops = (int64_t)h.hops;
  // the reference's loop runs at least once and stops at the limit
  int loops = h.max_iters > 1 ? h.max_iters : 1;
  out->loops = loops;
  out->found = h.not_found == 0;
  (void)variant; (void)looplimit;
  return PP_OK;
}

// part: optional piece of a chunk-walk launch (pp_push_direction_search_host); the caller then
// zeroes the counters once and gives every piece its own scheduler counter
struct ChunkPart { int begin, end; int* sched; };

pp_status do_search(pp_mesh* mesh, const PsView& view, int ps_nelems, const pp_search_args* a,
                    const double* dir, double distance, bool push, int push_from_orig,
                    pp_search_stats* stats_host, cudaStream_t s, const ChunkPart* part = nullptr) {
  PP_REQUIRE(mesh && a, "null argument");
  PP_REQUIRE(a->x_tgt && a->elem_ids, "x_tgt and elem_ids are required");
  PP_REQUIRE(a->stride >= view.capacity, "stride smaller than capacity");
  PP_REQUIRE(ps_nelems == mesh->nelems, "particle structure and mesh disagree on nelems");
  // labels of adjacency.tpp:609, adjacency.hpp:1152, :553
  PP_TIME(s, a->variant == PP_SEARCH_2D_LEGACY ? "pumipic search_2d"
             : a->variant == PP_SEARCH_3D ? "Search Mesh 3d" : "pumipic search_mesh");
  struct { int capacity; } ps_{view.capacity};
  auto* ps = &ps_;
  SearchParams p;
  p.ps = view;
  p.walk = mesh->walk;
  p.walk_bcc = mesh->walk_bcc;
  p.staged = g_staged_walk;
  p.xo = a->x_orig; p.xt = a->x_tgt; p.stride = a->stride;
  p.elem_ids = a->elem_ids; p.ids_empty = a->elem_ids_empty;
  p.inter_faces = a->inter_faces; p.inter_points = a->inter_points;
  p.looplimit = a->looplimit; p.tol = mesh->tol; p.nelems = mesh->nelems;
  p.unmoved_sq = unmoved_threshold(mesh->tol);
  p.counters = (SearchCounters*)mesh->stats_dev;
  p.dir = dir; p.distance = distance; p.xt_rw = const_cast<double*>(a->x_tgt);
  p.push_from_orig = push_from_orig;
  p.elem2sides = mesh->elem2sides; p.dual = mesh->dual; p.ndual = 0;
  p.chunk_begin = part ? part->begin : 0;
  if (p.chunk_begin < view.first_chunk) p.chunk_begin = view.first_chunk;   // leading empty chunks
  p.chunk_end = part ? part->end : view.nchunks;
  if (p.chunk_begin > p.chunk_end) p.chunk_begin = p.chunk_end;
  p.sched = part ? part->sched : &p.counters->next_chunk;
  if (!part) PP_CUDA(cudaMemsetAsync(mesh->stats_dev, 0, sizeof(SearchCounters), s));
  if (ps->capacity > 0) {
    switch (a->variant) {
      case PP_SEARCH_NEW:
        PP_REQUIRE(a->x_orig, "x_orig is required");
        if (a->require_intersection) {
          PP_REQUIRE(a->inter_faces && a->inter_points, "intersection outputs are required");
          if (mesh->dim == 3) launch<3, M_RAY>(p, push, s); else launch<2, M_RAY>(p, push, s);
        } else if (part || (p.staged >= 2 && a->elem_ids_empty && view.nchunks > 0 && view.C == 32 &&
                            view.chunk_start)) {
          if (mesh->dim == 3) PP_TRY((launch_walk_scs<3, false>(p, push, s)));
          else PP_TRY((launch_walk_scs<2, false>(p, push, s)));
        } else if (p.staged) {
          if (mesh->dim == 3) PP_TRY((launch_walk_bcc<3, false>(p, push, s)));
          else PP_TRY((launch_walk_bcc<2, false>(p, push, s)));
        } else {
          if (mesh->dim == 3) launch<3, M_BCC>(p, push, s); else launch<2, M_BCC>(p, push, s);
        }
        break;
      case PP_SEARCH_2D_LEGACY:
        PP_REQUIRE(mesh->dim == 2, "search_mesh_2d needs a 2D mesh");
        PP_REQUIRE(!push, "fused push is only available for the new search API");
        // elem_ids_empty: the caller passes a fresh array of -1 (test/pseudoXGCm.cpp:147-153), so every
        // particle starts in its row element and the chunk walk applies
        if (p.staged >= 2 && a->elem_ids_empty && view.nchunks > 0 && view.C == 32 && view.chunk_start)
          PP_TRY((launch_walk_scs<2, true>(p, false, s)));
        else if (p.staged) PP_TRY((launch_walk_bcc<2, true>(p, false, s)));
        else launch<2, M_LEG2D>(p, false, s);
        break;
      case PP_SEARCH_3D_LEGACY: {
        PP_REQUIRE(mesh->dim == 3, "legacy search_mesh needs a 3D mesh");
        PP_REQUIRE(!push, "fused push is only available for the new search API");
        PP_REQUIRE(a->x_orig && a->inter_faces && a->inter_points, "xpoints / xface are required");
        int nd = 0;
        PP_CUDA(cudaMemcpyAsync(&nd, mesh->dual_off + mesh->nelems, sizeof(int), cudaMemcpyDeviceToHost, s));
        PP_CUDA(cudaStreamSynchronize(s));
        p.ndual = nd;
        launch<3, M_LEG3D>(p, false, s);
        break;
      }
      case PP_SEARCH_3D: {
        PP_REQUIRE(mesh->dim == 3, "search_mesh_3d needs a 3D mesh");
        PP_REQUIRE(!push, "fused push is only available for the new search API");
        PP_REQUIRE(a->x_orig && a->inter_faces && a->inter_points, "xpoints / xface are required");
        int nd = 0;
        PP_CUDA(cudaMemcpyAsync(&nd, mesh->dual_off + mesh->nelems, sizeof(int), cudaMemcpyDeviceToHost, s));
        PP_CUDA(cudaStreamSynchronize(s));
        p.ndual = nd;
        launch<3, M_S3D>(p, false, s);
        break;
      }
      default:
        PP_REQUIRE(false, "unknown search variant");
    }
    PP_KERNEL_CHECK();
  }
  if (stats_host) PP_TRY(read_stats(mesh, a->variant, a->looplimit, stats_host, s));
  return PP_OK;
}

}  // namespace

// internal entry for searches over a flat list of points (gyro ring map)
pp_status pp_search_view(pp_mesh* mesh, const PsView& view, const pp_search_args* args,
                         pp_search_stats* stats_host, cudaStream_t s) {
  return do_search(mesh, view, mesh->nelems, args, nullptr, 0.0, false, 0, stats_host, s);
}

extern "C" pp_status pp_search_mesh(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                    pp_search_stats* stats_host, pp_stream stream) {
  PP_REQUIRE(ps, "null particle structure");
  return do_search(mesh, ps->view(), ps->nelems, args, nullptr, 0.0, false, 0, stats_host,
                   (cudaStream_t)stream);
}

extern "C" void pp_search_set_l2_window(double fraction) { g_l2_window = fraction > 0 ? (fraction < 1 ? fraction : 1.0) : 0.0; }

extern "C" void pp_search_set_staged(int32_t on) { g_staged_walk = on < 0 ? 0 : (on > 2 ? 2 : on); }

extern "C" pp_status pp_search_last_stats(pp_mesh* mesh, pp_search_stats* stats_host,
                                          pp_stream stream) {
  PP_REQUIRE(mesh && stats_host, "null argument");
  return read_stats(mesh, 0, 0, stats_host, (cudaStream_t)stream);
}

extern "C" pp_status pp_push_direction_search(pp_mesh* mesh, pp_ps* ps, const double* dir,
                                              double distance, int32_t push_from_orig,
                                              const pp_search_args* args,
                                              pp_search_stats* stats_host, pp_stream stream) {
  PP_REQUIRE(dir, "null direction array");
  PP_REQUIRE(args && args->variant == PP_SEARCH_NEW, "fused push needs PP_SEARCH_NEW");
  PP_REQUIRE(args->x_orig, "x_orig is required");
  PP_REQUIRE(ps, "null particle structure");
  return do_search(mesh, ps->view(), ps->nelems, args, dir, distance, true, push_from_orig ? 1 : 0,
                   stats_host, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Host-buffer variant: particle columns live in (pinned) host memory.  The slot range is cut into
// pieces at chunk boundaries; H2D copies, the chunk-walk kernel and D2H copies of successive
// pieces run on three streams, so the PCIe transfers in both directions overlap the kernel.
// ------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxParts = 64;
struct HostPipe {
  double *x = nullptr, *dir = nullptr, *xt = nullptr;
  int* ids = nullptr;
  int* sched = nullptr;            // [kMaxParts]
  long stride = 0;
  const void* dir_owner = nullptr; // structure whose direction column hp->dir holds (h_dir == NULL reuses it)
  int dir_capacity = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t e_begin = nullptr, e_end = nullptr, e_in[kMaxParts], e_k[kMaxParts];
  std::vector<int> chunk_start;
};

pp_status hostpipe_get(pp_mesh* mesh, long stride, HostPipe** out) {
  HostPipe* hp = (HostPipe*)mesh->hostpipe;
  if (!hp) {
    hp = new HostPipe();
    PP_CUDA(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
    PP_CUDA(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
    PP_CUDA(cudaEventCreateWithFlags(&hp->e_begin, cudaEventDisableTiming));
    PP_CUDA(cudaEventCreateWithFlags(&hp->e_end, cudaEventDisableTiming));
    for (int i = 0; i < kMaxParts; ++i) {
      PP_CUDA(cudaEventCreateWithFlags(&hp->e_in[i], cudaEventDisableTiming));
      PP_CUDA(cudaEventCreateWithFlags(&hp->e_k[i], cudaEventDisableTiming));
    }
    PP_CUDA(cudaMalloc((void**)&hp->sched, kMaxParts * sizeof(int)));
    mesh->hostpipe = hp;
  }
  if (hp->stride < stride) {
    cudaFree(hp->x); cudaFree(hp->dir); cudaFree(hp->xt); cudaFree(hp->ids);
    hp->x = hp->dir = hp->xt = nullptr; hp->ids = nullptr; hp->stride = 0;
    PP_CUDA(cudaMalloc((void**)&hp->x, 3 * stride * sizeof(double)));
    PP_CUDA(cudaMalloc((void**)&hp->dir, 3 * stride * sizeof(double)));
    PP_CUDA(cudaMalloc((void**)&hp->xt, 3 * stride * sizeof(double)));
    PP_CUDA(cudaMalloc((void**)&hp->ids, stride * sizeof(int)));
    hp->stride = stride;
    hp->dir_owner = nullptr;
  }
  *out = hp;
  return PP_OK;
}
}  // namespace

void pp_hostpipe_destroy(pp_mesh* mesh) {
  HostPipe* hp = (HostPipe*)mesh->hostpipe;
  if (!hp) return;
  cudaFree(hp->x); cudaFree(hp->dir); cudaFree(hp->xt); cudaFree(hp->ids); cudaFree(hp->sched);
  cudaStreamDestroy(hp->s_in); cudaStreamDestroy(hp->s_out);
  cudaEventDestroy(hp->e_begin); cudaEventDestroy(hp->e_end);
  for (int i = 0; i < kMaxParts; ++i) { cudaEventDestroy(hp->e_in[i]); cudaEventDestroy(hp->e_k[i]); }
  delete hp;
  mesh->hostpipe = nullptr;
}

extern "C" pp_status pp_push_direction_search_host(pp_mesh* mesh, pp_ps* ps, const double* h_x_orig,
                                                   const double* h_dir, double* h_x_tgt,
                                                   int32_t* h_elem_ids, int64_t stride,
                                                   double distance, int32_t looplimit,
                                                   int32_t nparts, pp_search_stats* stats_host,
                                                   pp_stream stream) {
  PP_REQUIRE(mesh && ps && h_x_orig && h_x_tgt && h_elem_ids, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  PP_REQUIRE(ps->nelems == mesh->nelems, "particle structure and mesh disagree on nelems");
  cudaStream_t s = (cudaStream_t)stream;
  const PsView view = ps->view();
  const bool chunked = view.nchunks > 0 && view.C == 32 && view.chunk_start;
  if (nparts < 1) nparts = 8;
  if (nparts > kMaxParts) nparts = kMaxParts;
  if (!chunked || nparts > view.nchunks) nparts = 1;
  HostPipe* hp;
  PP_TRY(hostpipe_get(mesh, stride, &hp));
  const long dstride = hp->stride;
  const int cap = ps->capacity;
  // h_dir == NULL: the direction column uploaded by the previous call for this structure stays
  PP_REQUIRE(h_dir || (hp->dir_owner == (const void*)ps && hp->dir_capacity == cap),
             "h_dir may only be NULL after a call that uploaded the directions of this structure");
  if (h_dir) { hp->dir_owner = ps; hp->dir_capacity = cap; }
  pp_search_args a;
  a.variant = PP_SEARCH_NEW;
  a.x_orig = hp->x; a.x_tgt = hp->xt; a.stride = dstride;
  a.elem_ids = hp->ids; a.elem_ids_empty = 1; a.require_intersection = 0;
  a.inter_faces = nullptr; a.inter_points = nullptr; a.looplimit = looplimit;
  // piece boundaries in slots
  std::vector<int> cbeg(nparts + 1, 0), sbeg(nparts + 1, 0);
  if (chunked) {
    hp->chunk_start.resize(view.nchunks + 1);
    PP_CUDA(cudaMemcpyAsync(hp->chunk_start.data(), view.chunk_start, sizeof(int) * (view.nchunks + 1),
                            cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    // equal slot counts per piece (chunk widths differ)
    int c = 0;
    for (int i = 1; i < nparts; ++i) {
      const long target = (long)cap * i / nparts;
      while (c < view.nchunks && hp->chunk_start[c] < target) ++c;
      cbeg[i] = c; sbeg[i] = hp->chunk_start[c];
    }
    cbeg[nparts] = view.nchunks;
  }
  sbeg[nparts] = cap;
  PP_CUDA(cudaMemsetAsync(mesh->stats_dev, 0, sizeof(SearchCounters), s));
  PP_CUDA(cudaMemsetAsync(hp->sched, 0, kMaxParts * sizeof(int), s));
  PP_CUDA(cudaEventRecord(hp->e_begin, s));
  PP_CUDA(cudaStreamWaitEvent(hp->s_in, hp->e_begin, 0));
  PP_CUDA(cudaStreamWaitEvent(hp->s_out, hp->e_begin, 0));
  for (int i = 0; i < nparts; ++i) {
    const long lo = sbeg[i], n = sbeg[i + 1] - sbeg[i];
    if (n > 0)
      for (int k = 0; k < 3; ++k) {
        PP_CUDA(cudaMemcpyAsync(hp->x + k * dstride + lo, h_x_orig + k * stride + lo, n * sizeof(double),
                                cudaMemcpyHostToDevice, hp->s_in));
        if (h_dir)
          PP_CUDA(cudaMemcpyAsync(hp->dir + k * dstride + lo, h_dir + k * stride + lo, n * sizeof(double),
                                  cudaMemcpyHostToDevice, hp->s_in));
      }
    PP_CUDA(cudaEventRecord(hp->e_in[i], hp->s_in));
  }
  for (int i = 0; i < nparts; ++i) {
    PP_CUDA(cudaStreamWaitEvent(s, hp->e_in[i], 0));
    if (chunked) {
      ChunkPart part{cbeg[i], cbeg[i + 1], hp->sched + i};
      if (part.end > part.begin)
        PP_TRY(do_search(mesh, view, ps->nelems, &a, hp->dir, distance, true, 1, nullptr, s, &part));
    } else {
      PP_TRY(do_search(mesh, view, ps->nelems, &a, hp->dir, distance, true, 1, nullptr, s));
    }
    PP_CUDA(cudaEventRecord(hp->e_k[i], s));
    PP_CUDA(cudaStreamWaitEvent(hp->s_out, hp->e_k[i], 0));
    const long lo = sbeg[i], n = sbeg[i + 1] - sbeg[i];
    if (n > 0) {
      for (int k = 0; k < 3; ++k)
        PP_CUDA(cudaMemcpyAsync(h_x_tgt + k * stride + lo, hp->xt + k * dstride + lo, n * sizeof(double),
                                cudaMemcpyDeviceToHost, hp->s_out));
      PP_CUDA(cudaMemcpyAsync(h_elem_ids + lo, hp->ids + lo, n * sizeof(int), cudaMemcpyDeviceToHost,
                              hp->s_out));
    }
  }
  PP_CUDA(cudaEventRecord(hp->e_end, hp->s_out));
  PP_CUDA(cudaStreamWaitEvent(s, hp->e_end, 0));
  if (stats_host) PP_TRY(read_stats(mesh, PP_SEARCH_NEW, looplimit, stats_host, s));
  return PP_OK;
}

// ==========================================================================================
// Stepped walk: the phases of trace_particle_through_mesh (adjacency.tpp:461-640) as separate
// calls, for applications that supply their own per-iteration handler (the `Func` argument,
// called on the host between find_exit_face and set_new_element with the device arrays
// elem_ids, inter_faces, lastExit, inter_points, ptcl_done).  One kernel per phase over all slots
// like the reference -- the fused pp_search_mesh is the fast path whenever the handler is the
// stock RemoveParticleOnGeometricModelExit; with that handler both give identical arrays.
// ==========================================================================================
namespace {
struct TraceParams {
  PsView ps;
  const void* walk;
  const int* elem2sides;
  const int8_t* exposed;
  const int* side2elem;
  const double* xo;
  const double* xt;
  long stride;
  int* elem_ids;
  int ids_empty;
  int* inter_faces;
  double* inter_points;
  int require_x;
  int* done;
  int* last_exit;
  double tol;
  int* counter;
};

// setInitial :504-522, finishUnmoved :525-533, initializeIntersection :535-549,
// check_initial_parents :73-145
template <int DIM>
__global__ void k_trace_begin(TraceParams p) {
  using Rec = typename std::conditional<DIM == 3, Tet, Tri>::type;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.ps.capacity) return;
  int erow;
  const bool mask = pp_slot_lookup(p.ps, s, erow);
  int done = 0, E;
  if (p.ids_empty) {
    E = mask ? erow : -1;
    if (!mask) done = 1;
  } else {
    E = p.elem_ids[s];
    if ((mask && E == -1) || !mask) done = 1;
  }
  d3 org = {0, 0, 0}, tgt = {0, 0, 0};
  if (mask) {
    org = {p.xo[s], p.xo[p.stride + s], p.xo[2 * p.stride + s]};
    tgt = {p.xt[s], p.xt[p.stride + s], p.xt[2 * p.stride + s]};
    if (norm3(tgt - org) < p.tol) done = 1;
  }
  if (p.require_x) {
    p.inter_faces[s] = -1;
#pragma unroll
    for (int i = 0; i < DIM; ++i) p.inter_points[(long)DIM * s + i] = 0.0;
  }
  if (mask && !done) {
    Rec rec;
    load_rec(p.walk, E, rec);
    bool inside;
    if constexpr (DIM == 3) {
      double b[4];
      bcc_tet(rec, org, b);
      inside = all_positive<4>(b, p.tol);
    } else {
      double b[3];
      bcc_tri(rec, d2{org.x, org.y}, b);
      inside = all_positive<3>(b, p.tol);
    }
    if (!inside) { atomicAdd(p.counter, 1); E = -1; done = 1; }
  }
  if (p.ids_empty || mask) p.elem_ids[s] = E;
  p.done[s] = done;
  p.last_exit[s] = -1;
}

// find_exit_face :232-364
template <int DIM, bool BCC>
__global__ void k_trace_find_exit(TraceParams p) {
  using Rec = typename std::conditional<DIM == 3, Tet, Tri>::type;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.ps.capacity) return;
  const bool mask = (__ldg(p.ps.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (!mask || p.done[s]) return;
  const int E = p.elem_ids[s];
  Rec rec;
  load_rec(p.walk, E, rec);
  const d3 tgt = {p.xt[s], p.xt[p.stride + s], p.xt[2 * p.stride + s]};
  const int* sides = p.elem2sides + (long)(DIM + 1) * E;
  if constexpr (BCC) {
    int f;
    bool done;
    if constexpr (DIM == 3) {
      double b[4];
      bcc_tet(rec, tgt, b);
      done = all_positive<4>(b, kEps);
      f = min_index4(b);
    } else {
      double b[3];
      bcc_tri(rec, d2{tgt.x, tgt.y}, b);
      done = all_positive<3>(b, kEps);
      f = min3(b);
    }
    p.done[s] = done;
    p.last_exit[s] = __ldg(sides + f);
  } else {
    const d3 org = {p.xo[s], p.xo[p.stride + s], p.xo[2 * p.stride + s]};
    const int prev = p.last_exit[s];
    int le = -1;
    if constexpr (DIM == 3) {
      const d3 disp = tgt - org;
      const double seg = norm3(disp);
      const d3 dir = {disp.x / seg, disp.y / seg, disp.z / seg};
      double quality = -1;
      int best = -1;
#pragma unroll
      for (int fi = 0; fi < 4; ++fi) {
        const int F = __ldg(sides + fi);
        if (F == prev) continue;
        const unsigned code = (rec.codes >> (8 * fi)) & 0xffu;
        const d3 V0 = vert_of(rec, code & 3), V1 = vert_of(rec, (code >> 2) & 3),
                 V2 = vert_of(rec, (code >> 4) & 3);
        d3 xp;
        double dproj, closeness;
        const bool hit = ray_tri(V0, V1, V2, org, dir, p.tol, (code >> 6) & 1, xp, dproj, closeness);
        bool write = false;
        if (hit) { le = F; write = true; }
        if (dproj > -p.tol && (quality < 0 || closeness < quality) && le == -1) {
          quality = closeness; best = F; write = true;
        }
        if (write) {
          p.inter_points[3 * (long)s] = xp.x;
          p.inter_points[3 * (long)s + 1] = xp.y;
          p.inter_points[3 * (long)s + 2] = xp.z;
        }
      }
      if (le == -1) le = best;
    } else {
#pragma unroll
      for (int ei = 0; ei < 3; ++ei) {
        const int F = __ldg(sides + ei);
        if (F == prev) continue;
        const unsigned code = (rec.codes >> (8 * ei)) & 0xffu;
        d2 xp;
        const bool hit = line_edge(vert_of(rec, code & 3), vert_of(rec, (code >> 2) & 3),
                                   d2{org.x, org.y}, d2{tgt.x, tgt.y}, p.tol, (code >> 6) & 1, xp);
        if (hit) {
          le = F;
          p.inter_points[2 * (long)s] = xp.x;
          p.inter_points[2 * (long)s + 1] = xp.y;
        }
      }
    }
    p.last_exit[s] = le;
    p.done[s] = (le == -1);
  }
}

// check_model_intersection :366-387
__global__ void k_trace_model_exit(TraceParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.ps.capacity) return;
  const bool mask = (__ldg(p.ps.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (!mask || p.done[s]) return;
  const int bridge = p.last_exit[s];
  const bool ex = p.exposed[bridge] != 0;
  p.done[s] = ex;
  if (ex && p.require_x) p.inter_faces[s] = bridge;
  else if (ex) p.elem_ids[s] = -1;
}

// set_new_element :390-416
__global__ void k_trace_next_elem(TraceParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.ps.capacity) return;
  const bool mask = (__ldg(p.ps.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (!mask || p.done[s]) return;
  const int cur = p.elem_ids[s];
  const int bridge = p.last_exit[s];
  const int A = __ldg(p.side2elem + 2 * (long)bridge), B = __ldg(p.side2elem + 2 * (long)bridge + 1);
  p.elem_ids[s] = (A == cur) ? B : A;
}

// :568-573 counts the slots (masked or not) that are not done; :584-606 with `sweep` also removes
// the masked ones among them
__global__ void k_trace_pending(TraceParams p, int sweep) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool pending = false;
  if (s < p.ps.capacity) {
    const bool mask = (__ldg(p.ps.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
    if (sweep) {
      pending = mask && !p.done[s];
      if (pending) p.elem_ids[s] = -1;
    } else {
      pending = !p.done[s];
    }
  }
  const int n = __popc(__ballot_sync(0xffffffffu, pending));
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(p.counter, n);
}

pp_status trace_params(pp_mesh* mesh, pp_ps* ps, const pp_search_args* a, int32_t* done,
                       int32_t* last_exit, bool need_positions, TraceParams* p) {
  PP_REQUIRE(mesh && ps && a && done && last_exit, "null argument");
  PP_REQUIRE(a->variant == PP_SEARCH_NEW, "the stepped walk follows the new search API only");
  PP_REQUIRE(a->elem_ids, "elem_ids is required");
  PP_REQUIRE(ps->nelems == mesh->nelems, "particle structure and mesh disagree on nelems");
  if (need_positions) {
    PP_REQUIRE(a->x_orig && a->x_tgt, "x_orig and x_tgt are required");
    PP_REQUIRE(a->stride >= ps->capacity, "stride smaller than capacity");
  }
  if (a->require_intersection)
    PP_REQUIRE(a->inter_faces && a->inter_points, "intersection outputs are required");
  p->ps = ps->view();
  p->walk = mesh->walk;
  p->elem2sides = mesh->elem2sides;
  p->exposed = mesh->exposed;
  p->side2elem = mesh->side2elem;
  p->xo = a->x_orig; p->xt = a->x_tgt; p->stride = a->stride;
  p->elem_ids = a->elem_ids; p->ids_empty = a->elem_ids_empty;
  p->inter_faces = a->inter_faces; p->inter_points = a->inter_points;
  p->require_x = a->require_intersection;
  p->done = done; p->last_exit = last_exit;
  p->tol = mesh->tol;
  p->counter = (int*)mesh->stats_dev;   // scratch: the fused search's counters are not in use here
  return PP_OK;
}

pp_status trace_count(pp_mesh* mesh, int32_t* out, cudaStream_t s) {
  int h = 0;
  PP_CUDA(cudaMemcpyAsync(&h, mesh->stats_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  if (out) *out = h;
  return PP_OK;
}
constexpr int kTraceBlock = 128;
}  // namespace

extern "C" pp_status pp_trace_begin(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                    int32_t* ptcl_done, int32_t* last_exit, int32_t* not_in_elem,
                                    pp_stream stream) {
  TraceParams p;
  if (ps && ps->capacity == 0) { if (not_in_elem) *not_in_elem = 0; return PP_OK; }
  PP_TRY(trace_params(mesh, ps, args, ptcl_done, last_exit, true, &p));
  cudaStream_t s = (cudaStream_t)stream;
  PP_CUDA(cudaMemsetAsync(mesh->stats_dev, 0, sizeof(SearchCounters), s));
  const int grid = pp_div_up(p.ps.capacity, kTraceBlock);
  if (mesh->dim == 3) k_trace_begin<3><<<grid, kTraceBlock, 0, s>>>(p);
  else k_trace_begin<2><<<grid, kTraceBlock, 0, s>>>(p);
  PP_KERNEL_CHECK();
  return trace_count(mesh, not_in_elem, s);
}

extern "C" pp_status pp_trace_find_exit_face(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                             int32_t* ptcl_done, int32_t* last_exit,
                                             pp_stream stream) {
  TraceParams p;
  if (ps && ps->capacity == 0) return PP_OK;
  PP_TRY(trace_params(mesh, ps, args, ptcl_done, last_exit, true, &p));
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = pp_div_up(p.ps.capacity, kTraceBlock);
  const bool bcc = !args->require_intersection;   // useBcc = !requireIntersection (:490)
  if (mesh->dim == 3) {
    if (bcc) k_trace_find_exit<3, true><<<grid, kTraceBlock, 0, s>>>(p);
    else k_trace_find_exit<3, false><<<grid, kTraceBlock, 0, s>>>(p);
  } else {
    if (bcc) k_trace_find_exit<2, true><<<grid, kTraceBlock, 0, s>>>(p);
    else k_trace_find_exit<2, false><<<grid, kTraceBlock, 0, s>>>(p);
  }
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_trace_check_model_intersection(pp_mesh* mesh, pp_ps* ps,
                                                       const pp_search_args* args, int32_t* ptcl_done,
                                                       int32_t* last_exit, pp_stream stream) {
  TraceParams p;
  if (ps && ps->capacity == 0) return PP_OK;
  PP_TRY(trace_params(mesh, ps, args, ptcl_done, last_exit, false, &p));
  k_trace_model_exit<<<pp_div_up(p.ps.capacity, kTraceBlock), kTraceBlock, 0, (cudaStream_t)stream>>>(p);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_trace_set_new_element(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                              int32_t* ptcl_done, int32_t* last_exit,
                                              pp_stream stream) {
  TraceParams p;
  if (ps && ps->capacity == 0) return PP_OK;
  PP_TRY(trace_params(mesh, ps, args, ptcl_done, last_exit, false, &p));
  k_trace_next_elem<<<pp_div_up(p.ps.capacity, kTraceBlock), kTraceBlock, 0, (cudaStream_t)stream>>>(p);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_trace_pending(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                      int32_t* ptcl_done, int32_t* last_exit, int32_t remove_pending,
                                      int32_t* count, pp_stream stream) {
  TraceParams p;
  if (ps && ps->capacity == 0) { if (count) *count = 0; return PP_OK; }
  PP_TRY(trace_params(mesh, ps, args, ptcl_done, last_exit, false, &p));
  cudaStream_t s = (cudaStream_t)stream;
  PP_CUDA(cudaMemsetAsync(mesh->stats_dev, 0, sizeof(int), s));
  k_trace_pending<<<pp_div_up(p.ps.capacity, kTraceBlock), kTraceBlock, 0, s>>>(p, remove_pending);
  PP_KERNEL_CHECK();
  return trace_count(mesh, count, s);
}
