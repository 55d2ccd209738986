// pp_host_picpart.cpp -- host-side PICpart tags: which elements are safe on this rank, which
// other parts are buffered, and entity ownership.  Follows src/pumipic_part_construct.cpp:
// Mesh::Mesh(Input&) :73-114, bfsBufferLayers :409-441, bfsSafeInward :443-468 (BFS through
// the elements around every entity of Input::bridge_dim; 0 = vertices is the default) and
// defineOwners :304-323.  Setup-time code that the
// reference also runs on the host side of Omega_h; sub-mesh extraction for non-full PICparts is a
// "next" row.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "pp_host_internal.hpp"
#include "pumipic_b200.h"

void pp_set_error(const char* fmt, ...);

namespace pph {
Up build_up(int nverts, int nelems, int nv, const int32_t* ev) {
  Up u;
  u.off.assign((size_t)nverts + 1, 0);
  for (int64_t i = 0; i < (int64_t)nelems * nv; ++i) u.off[ev[i] + 1]++;
  for (int v = 0; v < nverts; ++v) u.off[v + 1] += u.off[v];
  u.val.resize(u.off[nverts]);
  std::vector<int> fill(nverts, 0);
  for (int e = 0; e < nelems; ++e)
    for (int k = 0; k < nv; ++k) {
      const int v = ev[(int64_t)e * nv + k];
      u.val[u.off[v] + fill[v]++] = e;
    }
  return u;
}
// ask_down(dim, bridge_dim) as a flat list: vertices and sides are stored, the edges of a tet are
// the union of the edges of its faces (order inside an element is irrelevant to the BFS)
bool elem_bridges(const HMesh& m, int bridge_dim, std::vector<int32_t>& out, int& per_elem) {
  const int dim = m.dim;
  if (bridge_dim < 0 || bridge_dim >= dim) return false;
  const int ne = m.nents[dim];
  if (bridge_dim == 0) {
    per_elem = dim + 1;
    out = m.verts[dim];
    return true;
  }
  if (bridge_dim == dim - 1) {
    per_elem = dim + 1;
    out = m.down[dim];
    return true;
  }
  per_elem = 6;   // edges of a tet
  out.assign((size_t)ne * 6, -1);
  for (int e = 0; e < ne; ++e) {
    int n = 0;
    int32_t* mine = out.data() + (size_t)e * 6;
    for (int f = 0; f < 4; ++f) {
      const int face = m.down[3][(size_t)e * 4 + f];
      for (int k = 0; k < 3; ++k) {
        const int edge = m.down[2][(size_t)face * 3 + k];
        bool seen = false;
        for (int j = 0; j < n; ++j) seen = seen || mine[j] == edge;
        if (!seen) {
          if (n == 6) return false;
          mine[n++] = edge;
        }
      }
    }
    if (n != 6) return false;
  }
  return true;
}
namespace {
// one BFS layer (part_construct.cpp:387-405): every element around a bridge that touches a
// visited element becomes visited
void bfs_layer(const Up& u, int nverts, const std::vector<int>& visited, std::vector<int>& next) {
  for (int b = 0; b < nverts; ++b) {
    bool here = false;
    for (int j = u.off[b]; j < u.off[b + 1]; ++j)
      if (visited[u.val[j]]) here = true;
    if (here)
      for (int j = u.off[b]; j < u.off[b + 1]; ++j) next[u.val[j]] = 1;
  }
}
}  // namespace

// Mesh::Mesh(Input&), part_construct.cpp:73-114: safe tag and buffered parts of one rank.
void picpart_tags(const Up& u, int nverts, int nelems, const int32_t* owner, int nranks, int rank,
                  int buffer_method, int safe_method, int buffer_layers, int safe_layers,
                  std::vector<int>& is_safe, std::vector<int>& has_part) {
  enum { FULL = 0, BFS = 1, MINIMUM = 2, NONE = 3 };
  if (buffer_method == NONE) buffer_method = MINIMUM;        // pumipic_input.cpp:96-100
  if (buffer_method == MINIMUM) buffer_layers = 0;
  if (safe_method == MINIMUM) safe_layers = 0;
  is_safe.assign((size_t)nelems, safe_method == FULL);
  has_part.assign((size_t)nranks, 1);
  const bool need_bfs = (safe_method != NONE && safe_method != FULL) || buffer_method != FULL;
  if (need_bfs) {
    // bfsBufferLayers
    std::vector<int> safe(nelems, 0), part(nranks, 0), visited(nelems), next(nelems);
    for (int e = 0; e < nelems; ++e) visited[e] = next[e] = safe[e] = (owner[e] == rank);
    part[rank] = 1;
    for (int i = 0; i < buffer_layers || i < safe_layers; ++i) {
      bfs_layer(u, nverts, visited, next);
      for (int e = 0; e < nelems; ++e) {
        visited[e] = next[e];
        if (i == safe_layers - 1) safe[e] = next[e];
        if (i < buffer_layers && visited[e]) part[owner[e]] = 1;
      }
    }
    if (safe_method == BFS || safe_method == MINIMUM) is_safe = safe;
    if (buffer_method == BFS || buffer_method == MINIMUM) has_part = part;
  }
  if (buffer_method == BFS && safe_method == FULL) {
    // bfsSafeInward: everything is safe except safe_layers layers next to unbuffered parts
    std::vector<int> visited(nelems), next(nelems);
    for (int e = 0; e < nelems; ++e) visited[e] = next[e] = !has_part[owner[e]];
    for (int i = 0; i < safe_layers; ++i) {
      bfs_layer(u, nverts, visited, next);
      visited = next;
    }
    for (int e = 0; e < nelems; ++e) is_safe[e] = !visited[e] || owner[e] == rank;
  }
}
}  // namespace pph

extern "C" pp_status pp_host_picpart_tags(int32_t dim, int32_t nverts, int32_t nelems,
                                          const int32_t* elem2verts, const int32_t* owner,
                                          int32_t nranks, int32_t rank, int32_t buffer_method,
                                          int32_t safe_method, int32_t buffer_layers,
                                          int32_t safe_layers, int32_t* safe_out,
                                          int32_t* has_part_out) {
  if (!(dim == 2 || dim == 3) || !elem2verts || !owner || !safe_out || !has_part_out ||
      nranks < 1 || rank < 0 || rank >= nranks) {
    pp_set_error("pp_host_picpart_tags: bad argument");
    return PP_ERR_INVALID;
  }
  const pph::Up u = pph::build_up(nverts, nelems, dim + 1, elem2verts);
  std::vector<int> is_safe, has_part;
  pph::picpart_tags(u, nverts, nelems, owner, nranks, rank, buffer_method, safe_method,
                    buffer_layers, safe_layers, is_safe, has_part);
  memcpy(safe_out, is_safe.data(), sizeof(int32_t) * nelems);
  memcpy(has_part_out, has_part.data(), sizeof(int32_t) * nranks);
  return PP_OK;
}

extern "C" pp_status pp_host_picpart_tags_bridged(int32_t nbridges, int32_t nelems,
                                                  int32_t bridges_per_elem,
                                                  const int32_t* elem2bridges, const int32_t* owner,
                                                  int32_t nranks, int32_t rank,
                                                  int32_t buffer_method, int32_t safe_method,
                                                  int32_t buffer_layers, int32_t safe_layers,
                                                  int32_t* safe_out, int32_t* has_part_out) {
  if (nbridges < 0 || nelems < 0 || bridges_per_elem < 1 || !elem2bridges || !owner || !safe_out ||
      !has_part_out || nranks < 1 || rank < 0 || rank >= nranks) {
    pp_set_error("pp_host_picpart_tags_bridged: bad argument");
    return PP_ERR_INVALID;
  }
  for (int64_t i = 0; i < (int64_t)nelems * bridges_per_elem; ++i)
    if (elem2bridges[i] < 0 || elem2bridges[i] >= nbridges) {
      pp_set_error("pp_host_picpart_tags_bridged: bridge entity %d outside [0,%d)", elem2bridges[i], nbridges);
      return PP_ERR_INVALID;
    }
  for (int e = 0; e < nelems; ++e)
    if (owner[e] < 0 || owner[e] >= nranks) {
      pp_set_error("pp_host_picpart_tags_bridged: element %d has owner %d outside [0,%d)", e, owner[e], nranks);
      return PP_ERR_INVALID;
    }
  const pph::Up u = pph::build_up(nbridges, nelems, bridges_per_elem, elem2bridges);
  std::vector<int> is_safe, has_part;
  pph::picpart_tags(u, nbridges, nelems, owner, nranks, rank, buffer_method, safe_method,
                    buffer_layers, safe_layers, is_safe, has_part);
  if (nelems) memcpy(safe_out, is_safe.data(), sizeof(int32_t) * (size_t)nelems);
  memcpy(has_part_out, has_part.data(), sizeof(int32_t) * (size_t)nranks);
  return PP_OK;
}

extern "C" pp_status pp_host_entity_owners(int32_t nents, int32_t nelems, int32_t ents_per_elem,
                                           const int32_t* elem2ents, const int32_t* elem_owner,
                                           int32_t nranks, int32_t* ent_owner_out) {
  if (!elem2ents || !elem_owner || !ent_owner_out || nents < 0) {
    pp_set_error("pp_host_entity_owners: bad argument");
    return PP_ERR_INVALID;
  }
  for (int i = 0; i < nents; ++i) ent_owner_out[i] = nranks;   // defineOwners: min over adjacent elements
  for (int e = 0; e < nelems; ++e)
    for (int k = 0; k < ents_per_elem; ++k) {
      const int x = elem2ents[(int64_t)e * ents_per_elem + k];
      if (elem_owner[e] < ent_owner_out[x]) ent_owner_out[x] = elem_owner[e];
    }
  return PP_OK;
}

// Sub-mesh extraction for a partially buffered PICpart (constructPICPart,
// part_construct.cpp:116-262): the elements whose owner's core is buffered here stay
// (setSafeEnts :467-489), entities keep their relative order (offset_scan of the keep flags,
// :182-195), vertices are the vertices of the kept elements, coordinates are gathered
// (gatherCoords :499-511).  Output arrays are malloc'd; free with pp_host_free.
extern "C" pp_status pp_host_picpart_extract(int32_t dim, int32_t nverts, int32_t nelems,
                                             const double* coords, const int32_t* elem2verts,
                                             const int32_t* owner, int32_t nranks,
                                             const int32_t* has_part, int32_t* nelems_out,
                                             int32_t** elem_l2g_out, int32_t* nverts_out,
                                             int32_t** vert_l2g_out, int32_t** elem2verts_out,
                                             double** coords_out) {
  if (!(dim == 2 || dim == 3) || !coords || !elem2verts || !owner || !has_part || !nelems_out ||
      !elem_l2g_out || !nverts_out || !vert_l2g_out || !elem2verts_out || !coords_out || nranks < 1) {
    pp_set_error("pp_host_picpart_extract: bad argument");
    return PP_ERR_INVALID;
  }
  const int nv = dim + 1;
  std::vector<int> vkeep(nverts, 0);
  int ne_l = 0;
  for (int e = 0; e < nelems; ++e) {
    if (owner[e] < 0 || owner[e] >= nranks) {
      pp_set_error("pp_host_picpart_extract: element %d has owner %d outside [0,%d)", e, owner[e], nranks);
      return PP_ERR_INVALID;
    }
    if (!has_part[owner[e]]) continue;
    ++ne_l;
    for (int k = 0; k < nv; ++k) vkeep[elem2verts[(int64_t)e * nv + k]] = 1;
  }
  std::vector<int> vnum(nverts, -1);
  int nv_l = 0;
  for (int v = 0; v < nverts; ++v)
    if (vkeep[v]) vnum[v] = nv_l++;
  int32_t* el2g = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne_l > 0 ? ne_l : 1));
  int32_t* vl2g = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nv_l > 0 ? nv_l : 1));
  int32_t* ev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne_l > 0 ? ne_l : 1) * nv);
  double* co = (double*)malloc(sizeof(double) * (size_t)(nv_l > 0 ? nv_l : 1) * dim);
  if (!el2g || !vl2g || !ev || !co) {
    free(el2g); free(vl2g); free(ev); free(co);
    pp_set_error("pp_host_picpart_extract: out of memory");
    return PP_ERR_NOMEM;
  }
  int j = 0;
  for (int e = 0; e < nelems; ++e) {
    if (!has_part[owner[e]]) continue;
    el2g[j] = e;
    for (int k = 0; k < nv; ++k) ev[(int64_t)j * nv + k] = vnum[elem2verts[(int64_t)e * nv + k]];
    ++j;
  }
  for (int v = 0; v < nverts; ++v)
    if (vkeep[v]) {
      vl2g[vnum[v]] = v;
      for (int d = 0; d < dim; ++d) co[(int64_t)vnum[v] * dim + d] = coords[(int64_t)v * dim + d];
    }
  *nelems_out = ne_l; *elem_l2g_out = el2g; *nverts_out = nv_l; *vert_l2g_out = vl2g;
  *elem2verts_out = ev; *coords_out = co;
  return PP_OK;
}
