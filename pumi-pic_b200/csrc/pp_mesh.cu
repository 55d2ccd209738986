// pp_mesh.cu -- mesh handle: device-side derivation of everything the reference recomputes
// on every search call (measure_elements_real, mark_exposed_sides, ask_up(dim-1,dim), ask_dual,
// compute_tolerance_from_area; src/pumipic_adjacency.tpp:489-501,621) plus the packed
// per-element walk records that let one hop of the adjacency walk be a single aligned gather.
#include <cub/cub.cuh>

#include "pp_internal.cuh"

void pp_hostpipe_destroy(pp_mesh* mesh);   // pp_search.cu

namespace {

constexpr int kBlock = 256;

__constant__ int c_tet_face[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};
__constant__ int c_face_map[8] = {2, 1, 1, 3, 2, 3, 0, 3};  // pumipic_utils.hpp:489-493

__global__ void k_fill_int(int* a, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// ask_up(dim-1, dim): a side of a conforming simplicial mesh has one or two elements
__global__ void k_side_minmax(const int* __restrict__ elem2sides, int nent, int* lo, int* hi,
                              int nv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  int e = i / nv;
  int s = elem2sides[i];
  atomicMin(lo + s, e);
  atomicMax(hi + s, e);
}

__global__ void k_vert_first(const int* __restrict__ ev, int nent, int nv, int* first) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nent) atomicMin(first + ev[i], i / nv);
}

__global__ void k_side_finalize(const int* __restrict__ lo, const int* __restrict__ hi,
                                int nsides, int* side2elem, int8_t* exposed, int* n_exposed) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool ex = false;
  if (s < nsides) {
    int a = lo[s], b = hi[s];
    ex = (a == b);
    side2elem[2 * s] = a;
    side2elem[2 * s + 1] = ex ? -1 : b;
    exposed[s] = ex;  // mark_exposed_sides
  }
  unsigned m = __ballot_sync(0xffffffffu, ex);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_exposed, __popc(m));
}

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_val(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// measure_elements_real: tet_volume_from_basis / triangle_area_from_basis of simplex_basis
__global__ void k_measure(const double* __restrict__ coords, const int* __restrict__ ev,
                          int nelems, int dim, double* measure, unsigned long long* min_key) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long key = ~0ull;
  if (e < nelems) {
    double m;
    if (dim == 2) {
      const int* v = ev + 3 * (long)e;
      d2 p0 = {coords[2 * (long)v[0]], coords[2 * (long)v[0] + 1]};
      d2 p1 = {coords[2 * (long)v[1]], coords[2 * (long)v[1] + 1]};
      d2 p2 = {coords[2 * (long)v[2]], coords[2 * (long)v[2] + 1]};
      m = cross2(p1 - p0, p2 - p0) / 2.0;
    } else {
      const int* v = ev + 4 * (long)e;
      d3 p[4];
      for (int k = 0; k < 4; ++k)
        p[k] = {coords[3 * (long)v[k]], coords[3 * (long)v[k] + 1], coords[3 * (long)v[k] + 2]};
      m = dot3(cross3(p[1] - p[0], p[2] - p[0]), p[3] - p[0]) / 6.0;
    }
    measure[e] = m;
    key = ordered_key(m);
  }
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key != ~0ull) atomicMin(min_key, key);
}

__global__ void k_dual_count(const int* __restrict__ elem2sides, const int8_t* __restrict__ exposed,
                             int nelems, int nv, int* cnt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  int c = 0;
  for (int k = 0; k < nv; ++k) c += !exposed[elem2sides[(long)e * nv + k]];
  cnt[e] = c;
}

// ask_dual + packed walk record of each element
template <int DIM>
__global__ void k_build_walk(const double* __restrict__ coords, const int* __restrict__ ev,
                             const int* __restrict__ e2s, const int* __restrict__ s2v,
                             const int* __restrict__ side2elem, const int* __restrict__ cls,
                             const double* __restrict__ measure, const int* __restrict__ dual_off,
                             int nelems, int* dual, void* walk_out, PPBccRec3* bcc_out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  constexpr int NV = DIM + 1;
  int tv[NV];
  for (int k = 0; k < NV; ++k) tv[k] = ev[(long)e * NV + k];
  int adj[NV];
  unsigned codes = 0;
  int dpos = dual_off[e];
  for (int f = 0; f < NV; ++f) {
    const int s = e2s[(long)e * NV + f];
    const int a = side2elem[2 * (long)s], b = side2elem[2 * (long)s + 1];
    if (b < 0) {
      adj[f] = -s - 1;
    } else {
      adj[f] = (a == e) ? b : a;
      dual[dpos++] = adj[f];
    }
    unsigned code = 0;
    int fv[DIM];
    for (int k = 0; k < DIM; ++k) {
      fv[k] = s2v[(long)s * DIM + k];
      int loc = 0;
      for (int j = 0; j < NV; ++j)
        if (tv[j] == fv[k]) loc = j;
      code |= (unsigned)loc << (2 * k);
    }
    if (DIM == 3) {
      const int m1 = c_face_map[2 * f], m2 = c_face_map[2 * f + 1];
      const int idx = (fv[0] == tv[m1]) ? 1 : (fv[1] == tv[m1]) ? 2 : 0;
      const bool flip = tv[m2] != fv[idx];                         // utils.hpp:501-507
      const bool lflip = !(fv[1] == tv[m1] && fv[DIM - 1] == tv[m2]);  // adjacency.hpp:662-664
      code |= (flip ? 1u : 0u) << 6;
      code |= (lflip ? 1u : 0u) << 7;
    } else {
      const int idx = (fv[0] == tv[0]) ? 1 : (fv[0] == tv[1]) ? 2 : 0;
      const bool flip = fv[1] != tv[idx];                          // utils.hpp:495-499
      code |= (flip ? 1u : 0u) << 6;
    }
    codes |= code << (8 * f);
  }
  if constexpr (DIM == 3) {
    PPTetRec r;
    for (int k = 0; k < 4; ++k)
      for (int i = 0; i < 3; ++i) r.c[3 * k + i] = coords[3 * (long)tv[k] + i];
    r.vol = measure[e];
    for (int f = 0; f < 4; ++f) r.adj[f] = adj[f];
    r.codes = codes;
    r.aux = -1;
    ((PPTetRec*)walk_out)[e] = r;
    PPBccRec3 b;
    for (int i = 0; i < 9; ++i) b.a[i] = r.c[i];
    for (int f = 0; f < 4; ++f) {   // barycentric_tet: cross(vac, vab), (a,b,c) = face template
      const double* pa = r.c + 3 * c_tet_face[f][0];
      const double* pb = r.c + 3 * c_tet_face[f][1];
      const double* pc = r.c + 3 * c_tet_face[f][2];
      const d3 vab = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
      const d3 vac = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
      const d3 n = cross3(vac, vab);
      b.n[3 * f] = n.x; b.n[3 * f + 1] = n.y; b.n[3 * f + 2] = n.z;
    }
    b.inv_vol = r.vol > 0 ? 1.0 / r.vol : -1.0;
    for (int f = 0; f < 4; ++f) b.adj[f] = adj[f];
    bcc_out[e] = b;
  } else {
    PPTriRec r;
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 2; ++i) r.c[2 * k + i] = coords[2 * (long)tv[k] + i];
    r.area = measure[e];
    for (int f = 0; f < 3; ++f) r.adj[f] = adj[f];
    r.codes = codes;
    r.cls = cls ? cls[e] : 0;
    r.aux = -1;
    r.pad[0] = r.pad[1] = 0;
    ((PPTriRec*)walk_out)[e] = r;
  }
}

__global__ void k_set_aux(void* walk, int dim, int nelems, const int* __restrict__ safe,
                          const int* __restrict__ owner, int* aux_out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  const int aux = safe[e] ? -1 : owner[e];
  aux_out[e] = aux;
  if (dim == 3) ((PPTetRec*)walk)[e].aux = aux;
  else ((PPTriRec*)walk)[e].aux = aux;
}

}  // namespace

extern "C" pp_status pp_mesh_create(const pp_mesh_desc* d, pp_stream stream_, pp_mesh** out) {
  PP_REQUIRE(d && out, "null argument");
  PP_REQUIRE(d->dim == 2 || d->dim == 3, "dim must be 2 or 3");
  PP_REQUIRE(d->nverts > 0 && d->nelems > 0 && d->nsides > 0, "empty mesh");
  PP_REQUIRE(d->coords && d->elem2verts && d->elem2sides && d->side2verts, "null mesh array");
  pp_runtime_init();
  cudaStream_t s = (cudaStream_t)stream_;
  const int dim = d->dim, nv = dim + 1, ne = d->nelems, ns = d->nsides;
  pp_mesh* m = new pp_mesh();
  memset(m, 0, sizeof(*m));
  m->dim = dim; m->nverts = d->nverts; m->nelems = ne; m->nsides = ns; m->self_rank = 0;
  PP_TRY(pp_dev_import(&m->coords, d->coords, (size_t)d->nverts * dim, d->memspace, s));
  PP_TRY(pp_dev_import(&m->elem2verts, d->elem2verts, (size_t)ne * nv, d->memspace, s));
  PP_TRY(pp_dev_import(&m->elem2sides, d->elem2sides, (size_t)ne * nv, d->memspace, s));
  PP_TRY(pp_dev_import(&m->side2verts, d->side2verts, (size_t)ns * dim, d->memspace, s));
  if (d->elem_class) PP_TRY(pp_dev_import(&m->elem_class, d->elem_class, (size_t)ne, d->memspace, s));

  int *lo, *hi, *cnt, *scal;
  unsigned long long* min_key;
  PP_TRY(pp_dev_alloc(&lo, ns, s));
  PP_TRY(pp_dev_alloc(&hi, ns, s));
  PP_TRY(pp_dev_alloc(&cnt, ne + 1, s));
  PP_TRY(pp_dev_alloc(&scal, 4, s));
  PP_TRY(pp_dev_alloc(&min_key, 1, s));
  PP_TRY(pp_dev_alloc(&m->side2elem, 2 * (size_t)ns, s));
  PP_TRY(pp_dev_alloc(&m->exposed, ns, s));
  PP_TRY(pp_dev_alloc(&m->measure, ne, s));
  PP_TRY(pp_dev_alloc(&m->dual_off, ne + 1, s));
  PP_TRY(pp_dev_alloc(&m->stats_dev, sizeof(SearchCounters) / sizeof(int), s));
  PP_CUDA(cudaMemsetAsync(scal, 0, 4 * sizeof(int), s));
  PP_CUDA(cudaMemsetAsync(min_key, 0xff, sizeof(unsigned long long), s));
  PP_CUDA(cudaMemsetAsync(m->stats_dev, 0, sizeof(SearchCounters), s));

  k_fill_int<<<pp_div_up(ns, kBlock), kBlock, 0, s>>>(lo, ns, 0x7fffffff);
  k_fill_int<<<pp_div_up(ns, kBlock), kBlock, 0, s>>>(hi, ns, -1);
  k_side_minmax<<<pp_div_up((long)ne * nv, kBlock), kBlock, 0, s>>>(m->elem2sides, ne * nv, lo, hi, nv);
  k_side_finalize<<<pp_div_up(ns, kBlock), kBlock, 0, s>>>(lo, hi, ns, m->side2elem, m->exposed, scal);
  k_measure<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(m->coords, m->elem2verts, ne, dim, m->measure, min_key);
  k_dual_count<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(m->elem2sides, m->exposed, ne, nv, cnt);
  PP_KERNEL_CHECK();
  PP_CUDA(cudaMemsetAsync(cnt + ne, 0, sizeof(int), s));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, m->dual_off, ne + 1, s);
  void* tmp;
  PP_TRY(pp_dev_alloc((char**)&tmp, tmp_bytes, s));
  PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, m->dual_off, ne + 1, s));
  int ndual = 0;
  PP_CUDA(cudaMemcpyAsync(&ndual, m->dual_off + ne, sizeof(int), cudaMemcpyDeviceToHost, s));
  unsigned long long hkey = 0;
  PP_CUDA(cudaMemcpyAsync(&hkey, min_key, sizeof(hkey), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaMemcpyAsync(&m->n_exposed, scal, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  {
    unsigned long long b = (hkey >> 63) ? (hkey & 0x7fffffffffffffffull) : ~hkey;
    memcpy(&m->min_measure, &b, sizeof(double));
  }
  // compute_tolerance_from_area (adjacency.tpp:419-428)
  const double t = 1e-15 / m->min_measure;
  m->tol = t > 1e-8 ? t : 1e-8;

  PP_TRY(pp_dev_alloc(&m->dual, (size_t)ndual + 1, s));
  const size_t rec = dim == 3 ? sizeof(PPTetRec) : sizeof(PPTriRec);
  PP_CUDA(cudaMallocAsync(&m->walk, rec * (size_t)ne, s));
  PP_TRY(pp_dev_alloc(&m->vert_first_elem, d->nverts, s));
  k_fill_int<<<pp_div_up(d->nverts, kBlock), kBlock, 0, s>>>(m->vert_first_elem, d->nverts, 0x7fffffff);
  k_vert_first<<<pp_div_up((long)ne * nv, kBlock), kBlock, 0, s>>>(m->elem2verts, ne * nv, nv, m->vert_first_elem);
  PP_TRY(pp_dev_alloc(&m->aux, ne, s));
  k_fill_int<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(m->aux, ne, -1);
  if (dim == 3) {
    PP_TRY(pp_dev_alloc(&m->walk_bcc, ne, s));
    k_build_walk<3><<<pp_div_up(ne, 128), 128, 0, s>>>(m->coords, m->elem2verts, m->elem2sides,
                                                       m->side2verts, m->side2elem, m->elem_class,
                                                       m->measure, m->dual_off, ne, m->dual, m->walk,
                                                       m->walk_bcc);
  } else {
    k_build_walk<2><<<pp_div_up(ne, 128), 128, 0, s>>>(m->coords, m->elem2verts, m->elem2sides,
                                                       m->side2verts, m->side2elem, m->elem_class,
                                                       m->measure, m->dual_off, ne, m->dual, m->walk,
                                                       nullptr);
  }
  PP_KERNEL_CHECK();
  pp_dev_free(lo, s); pp_dev_free(hi, s); pp_dev_free(cnt, s); pp_dev_free(scal, s);
  pp_dev_free(min_key, s); pp_dev_free((char*)tmp, s);
  PP_CUDA(cudaStreamSynchronize(s));
  *out = m;
  return PP_OK;
}

extern "C" pp_status pp_mesh_destroy(pp_mesh* m) {
  if (!m) return PP_OK;
  cudaFree(m->coords); cudaFree(m->elem2verts); cudaFree(m->elem2sides); cudaFree(m->side2verts);
  cudaFree(m->elem_class); cudaFree(m->measure); cudaFree(m->exposed); cudaFree(m->side2elem);
  cudaFree(m->dual_off); cudaFree(m->dual); cudaFree(m->safe); cudaFree(m->owner);
  cudaFree(m->walk); cudaFree(m->walk_bcc); cudaFree(m->aux); cudaFree(m->vert_first_elem); cudaFree(m->stats_dev);
  pp_hostpipe_destroy(m);
  delete m;
  return PP_OK;
}

extern "C" pp_status pp_mesh_get_info(const pp_mesh* m, pp_mesh_info* o) {
  PP_REQUIRE(m && o, "null argument");
  o->dim = m->dim; o->nverts = m->nverts; o->nelems = m->nelems; o->nsides = m->nsides;
  o->tol = m->tol; o->min_measure = m->min_measure; o->n_exposed_sides = m->n_exposed;
  o->walk_table_bytes = (int64_t)m->nelems *
                        (m->dim == 3 ? sizeof(PPTetRec) + sizeof(PPBccRec3) : sizeof(PPTriRec));
  return PP_OK;
}

extern "C" pp_status pp_mesh_get_arrays(const pp_mesh* m, pp_mesh_arrays* o) {
  PP_REQUIRE(m && o, "null argument");
  o->coords = m->coords; o->elem2verts = m->elem2verts; o->elem2sides = m->elem2sides;
  o->side2verts = m->side2verts; o->elem_class = m->elem_class; o->measure = m->measure;
  o->exposed = m->exposed; o->side2elem = m->side2elem; o->dual_off = m->dual_off;
  o->dual = m->dual;
  return PP_OK;
}

extern "C" pp_status pp_mesh_set_picpart(pp_mesh* m, const int32_t* safe, const int32_t* owner,
                                         int32_t self_rank, int32_t memspace, pp_stream stream_) {
  PP_REQUIRE(m && safe && owner, "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  if (m->safe) { cudaFree(m->safe); m->safe = nullptr; }
  if (m->owner) { cudaFree(m->owner); m->owner = nullptr; }
  PP_TRY(pp_dev_import(&m->safe, safe, (size_t)m->nelems, memspace, s));
  PP_TRY(pp_dev_import(&m->owner, owner, (size_t)m->nelems, memspace, s));
  m->self_rank = self_rank;
  k_set_aux<<<pp_div_up(m->nelems, kBlock), kBlock, 0, s>>>(m->walk, m->dim, m->nelems, m->safe, m->owner, m->aux);
  PP_KERNEL_CHECK();
  return PP_OK;
}
