// pp_host_mesh.cpp -- host-side mesh utilities (no CUDA): side derivation for simplicial meshes
// and the synthetic generators used by the benchmarks (Kuhn-split cube, triangulated plate).
// The reference gets all of this from Omega_h (not vendored); these helpers exist so that a
// caller without Omega_h can still hand pp_mesh_create a complete description.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <numeric>
#include <vector>

#include "pumipic_b200.h"

void pp_set_error(const char* fmt, ...);

namespace {
const int kTetFace[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};  // Omega_h simplex templates
const int kTriEdge[3][2] = {{0, 1}, {1, 2}, {2, 0}};

struct SideKey {
  int32_t v[3];
  int64_t flat;
};
}  // namespace

extern "C" pp_status pp_host_derive_sides(int32_t dim, int32_t nelems, const int32_t* ev,
                                          int32_t* nsides_out, int32_t** e2s_out,
                                          int32_t** s2v_out) {
  if (!(dim == 2 || dim == 3) || nelems <= 0 || !ev || !nsides_out || !e2s_out || !s2v_out) {
    pp_set_error("pp_host_derive_sides: bad argument");
    return PP_ERR_INVALID;
  }
  const int nv = dim + 1;
  const int64_t n = (int64_t)nelems * nv;
  std::vector<SideKey> keys((size_t)n);
  auto side_verts = [&](int64_t flat, int32_t out[3]) {
    const int64_t e = flat / nv;
    const int k = (int)(flat % nv);
    for (int i = 0; i < dim; ++i) {
      const int loc = dim == 3 ? kTetFace[k][i] : kTriEdge[k][i];
      out[i] = ev[e * nv + loc];
    }
    if (dim == 2) out[2] = -1;
  };
  for (int64_t f = 0; f < n; ++f) {
    int32_t v[3];
    side_verts(f, v);
    std::sort(v, v + dim);
    keys[(size_t)f] = {{v[0], v[1], v[2]}, f};
  }
  std::sort(keys.begin(), keys.end(), [](const SideKey& a, const SideKey& b) {
    if (a.v[0] != b.v[0]) return a.v[0] < b.v[0];
    if (a.v[1] != b.v[1]) return a.v[1] < b.v[1];
    if (a.v[2] != b.v[2]) return a.v[2] < b.v[2];
    return a.flat < b.flat;
  });
  int32_t* e2s = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  std::vector<int64_t> first;
  first.reserve((size_t)n / 2 + 16);
  int32_t ns = -1;
  for (int64_t i = 0; i < n; ++i) {
    const SideKey& k = keys[(size_t)i];
    if (i == 0 || k.v[0] != keys[(size_t)i - 1].v[0] || k.v[1] != keys[(size_t)i - 1].v[1] ||
        k.v[2] != keys[(size_t)i - 1].v[2]) {
      ++ns;
      first.push_back(k.flat);  // lowest flat index of the group: sorted ascending within a key
    }
    e2s[k.flat] = ns;
  }
  ++ns;
  int32_t* s2v = (int32_t*)malloc(sizeof(int32_t) * (size_t)ns * dim);
  for (int32_t s = 0; s < ns; ++s) {
    int32_t v[3];
    side_verts(first[(size_t)s], v);
    for (int i = 0; i < dim; ++i) s2v[(size_t)s * dim + i] = v[i];
  }
  *nsides_out = ns;
  *e2s_out = e2s;
  *s2v_out = s2v;
  return PP_OK;
}

extern "C" pp_status pp_host_kuhn_cube(int32_t n, double length, int32_t* nverts_out,
                                       double** coords_out, int32_t* nelems_out,
                                       int32_t** ev_out) {
  if (n <= 0 || !nverts_out || !coords_out || !nelems_out || !ev_out) {
    pp_set_error("pp_host_kuhn_cube: bad argument");
    return PP_ERR_INVALID;
  }
  const int64_t s = n + 1;
  const int64_t nverts = s * s * s, nelems = 6 * (int64_t)n * n * n;
  double* coords = (double*)malloc(sizeof(double) * (size_t)nverts * 3);
  int32_t* ev = (int32_t*)malloc(sizeof(int32_t) * (size_t)nelems * 4);
  for (int64_t k = 0; k < s; ++k)
    for (int64_t j = 0; j < s; ++j)
      for (int64_t i = 0; i < s; ++i) {
        const int64_t v = i + s * (j + s * k);
        coords[3 * v + 0] = (double)i * (length / n);
        coords[3 * v + 1] = (double)j * (length / n);
        coords[3 * v + 2] = (double)k * (length / n);
      }
  static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  static const int odd[6] = {0, 1, 1, 0, 0, 1};
  int64_t e = 0;
  for (int64_t k = 0; k < n; ++k)
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i < n; ++i)
        for (int p = 0; p < 6; ++p) {
          int64_t c[3] = {i, j, k};
          int32_t v[4];
          v[0] = (int32_t)(c[0] + s * (c[1] + s * c[2]));
          for (int t = 0; t < 3; ++t) {
            c[perms[p][t]] += 1;
            v[t + 1] = (int32_t)(c[0] + s * (c[1] + s * c[2]));
          }
          if (odd[p]) std::swap(v[1], v[2]);  // keep tet_volume_from_basis positive
          memcpy(ev + 4 * e, v, sizeof(v));
          ++e;
        }
  *nverts_out = (int32_t)nverts;
  *coords_out = coords;
  *nelems_out = (int32_t)nelems;
  *ev_out = ev;
  return PP_OK;
}

extern "C" pp_status pp_host_plate(int32_t n, double length, int32_t* nverts_out,
                                   double** coords_out, int32_t* nelems_out, int32_t** ev_out) {
  if (n <= 0 || !nverts_out || !coords_out || !nelems_out || !ev_out) {
    pp_set_error("pp_host_plate: bad argument");
    return PP_ERR_INVALID;
  }
  const int64_t s = n + 1;
  const int64_t nverts = s * s, nelems = 2 * (int64_t)n * n;
  double* coords = (double*)malloc(sizeof(double) * (size_t)nverts * 2);
  int32_t* ev = (int32_t*)malloc(sizeof(int32_t) * (size_t)nelems * 3);
  for (int64_t j = 0; j < s; ++j)
    for (int64_t i = 0; i < s; ++i) {
      coords[2 * (i + s * j) + 0] = (double)i * (length / n);
      coords[2 * (i + s * j) + 1] = (double)j * (length / n);
    }
  int64_t e = 0;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const int32_t v00 = (int32_t)(i + s * j), v10 = v00 + 1, v01 = (int32_t)(v00 + s), v11 = v01 + 1;
      const int32_t t0[3] = {v00, v10, v11}, t1[3] = {v00, v11, v01};
      memcpy(ev + 3 * e, t0, sizeof(t0)); ++e;
      memcpy(ev + 3 * e, t1, sizeof(t1)); ++e;
    }
  *nverts_out = (int32_t)nverts;
  *coords_out = coords;
  *nelems_out = (int32_t)nelems;
  *ev_out = ev;
  return PP_OK;
}

extern "C" void pp_host_free(void* p) { free(p); }
