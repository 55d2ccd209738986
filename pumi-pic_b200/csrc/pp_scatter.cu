// pp_scatter.cu -- XGC-like 2D path: elliptical push, gyro ring mapping, gyro-averaged charge
// scatter, setUnsafeProcs.
//
// Replaces test/ellipticalPush.hpp:10-70, test/gyroScatter.hpp:25-258 and
// src/pumipic_ptcl_ops.hpp:33-53.  The reference's accumulateToRings issues 6 fp64 atomics per
// PARTICLE onto the 3 vertices of its element although every addend is the constant 1.0
// (gyroScatter.hpp:182-204); here the per-element particle count comes from the structure itself
// and each ELEMENT issues the 6 atomics once with its count.  All values are exact integers in
// fp64, so the result is bit-identical in any order.
#include "pp_internal.cuh"

pp_status pp_search_view(pp_mesh* mesh, const PsView& view, const pp_search_args* args,
                         pp_search_stats* stats_host, cudaStream_t s);

namespace {
constexpr int kBlock = 256;

__global__ void k_elliptical_setup(PsView v, const double* __restrict__ x, long stride, float* b,
                                   float* phi, double h, double k, double d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  if (!((__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u)) return;
  const double w = x[s], z = x[stride + s];
  const double ph = atan2(d * (z - k), w - h);
  const double bb = (z - k) / sin(ph);
  phi[s] = (float)ph;
  b[s] = (float)bb;
}

__global__ void k_elliptical_push(PsView v, const int* __restrict__ class_ids, double* xt, long stride,
                                  const float* __restrict__ b, float* phi, double h, double k,
                                  double d, double deg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  int e;
  if (!pp_slot_lookup(v, s, e)) return;
  const int cls = class_ids[e];
  const double centerFactor = cls == 1 ? 0.01 : 1.0;
  const double distByClass = centerFactor * (double)1.0 / cls;
  const double degP = deg * distByClass;
  const float ph = phi[s];
  const float bb = b[s];
  const double a = bb * d;
  const double rad = ph + degP * 3.14159265358979323846 / 180.0;
  xt[s] = a * cos(rad) + h;
  xt[stride + s] = bb * sin(rad) + k;
  phi[s] = (float)rad;
}

__global__ void k_set_unsafe(PsView v, const int* __restrict__ elems, const int* __restrict__ aux,
                             int self, int* new_elems, int* new_procs) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  const int e = elems[s];
  int proc = self;
  if (m && e != -1) {
    const int o = __ldg(aux + e);   // owner if the element is not safe, else -1
    if (o >= 0) proc = o;
  }
  new_elems[s] = e;
  new_procs[s] = proc;
}

// particles per element of a flat structure (DPS): warp-aggregated histogram
__global__ void k_count_flat(PsView v, int* cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int e = -1;
  if (s < v.capacity && ((__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u)) e = v.slot_elem[s];
  const unsigned grp = __match_any_sync(0xffffffffu, e);
  if (e >= 0 && (threadIdx.x & 31) == (__ffs(grp) - 1)) atomicAdd(cnt + e, __popc(grp));
}

// accumulateToRings, one thread per element
__global__ void k_rings(const int* __restrict__ ev, int ne, const int* __restrict__ cnt,
                        const int* __restrict__ row_ppe, const int* __restrict__ elem2row,
                        const int* __restrict__ csr_off, int gnr, int ringDown, int ringUp,
                        double* ring_accum) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int c;
  if (cnt) c = cnt[e];
  else if (csr_off) c = csr_off[e + 1] - csr_off[e];
  else c = row_ppe[elem2row[e]];
  if (c <= 0) return;
  const double w = (double)c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long v = ev[3 * (long)e + i];
    atomicAdd(ring_accum + v * gnr + ringUp, w);
    atomicAdd(ring_accum + v * gnr + ringDown, w);
  }
}

// scatterToMappedVerts (gyroScatter.hpp:207-224), one thread per (vertex, ring, point)
__global__ void k_scatter_mapped(const double* __restrict__ ring_accum, const int* __restrict__ v2v,
                                 long npts, int gnr, int gppr, double* scatter_w) {
  const long id = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (id >= npts) return;
  const long vr = id / gppr;            // v*gnr + ring
  const double val = ring_accum[vr] / gppr;
  if (val == 0.0) return;               // adding zero changes nothing
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int mv = v2v[3 * id + k];
    if (mv >= 0) atomicAdd(scatter_w + mv, val);
  }
}

__global__ void k_ring_points(const double* __restrict__ coords, const int* __restrict__ first_elem,
                              long npts, int gnr, int gppr, double rmax, double theta, double* tgt,
                              int* start, uint32_t* mask) {
  const long id = blockIdx.x * (long)blockDim.x + threadIdx.x;
  bool on = false;
  if (id < npts) {
    const int point_id = (int)(id % gppr);
    const long id2 = id / gppr;
    const int ring_id = (int)(id2 % gnr);
    const long vert_id = id2 / gnr;
    const double radius = rmax * (ring_id + 1) / gnr;
    const double deg = theta + (((double)point_id) / gppr * 360);
    const double rad = deg * (3.14159265358979323846 / 180);
    tgt[id] = coords[2 * vert_id] + radius * cos(rad);
    tgt[npts + id] = coords[2 * vert_id + 1] + radius * sin(rad);
    tgt[2 * npts + id] = 0;
    start[id] = first_elem[vert_id];
    on = true;
  }
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && id - (id & 31) < npts + 31) mask[id >> 5] = m;
}

__global__ void k_ring_map(const int* __restrict__ elem_ids, const int* __restrict__ ev, long npts, int* map) {
  const long id = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (id >= npts) return;
  const int parent = elem_ids[id];
#pragma unroll
  for (int i = 0; i < 3; ++i) map[3 * id + i] = parent >= 0 ? ev[3 * (long)parent + i] : -1;
}

__global__ void k_interleave(const double* __restrict__ a, const double* __restrict__ b, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[2 * i] = a[i];
  out[2 * i + 1] = b[i];
}
}  // namespace

extern "C" pp_status pp_push_elliptical_setup(pp_ps* ps, const double* x, int64_t stride, float* b,
                                              float* phi, double h, double k, double d,
                                              pp_stream stream) {
  PP_REQUIRE(ps && x && b && phi, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_elliptical_setup<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, stride, b, phi, h, k, d);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_push_elliptical(pp_mesh* mesh, pp_ps* ps, double* xtgt, int64_t stride,
                                        const float* b, float* phi, double h, double k, double d,
                                        double deg, pp_stream stream) {
  PP_REQUIRE(mesh && ps && xtgt && b && phi, "null argument");
  PP_REQUIRE(mesh->elem_class, "the mesh was created without element class ids");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_elliptical_push<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), mesh->elem_class, xtgt, stride, b, phi, h, k, d, deg);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_set_unsafe_procs(pp_mesh* mesh, pp_ps* ps, const int32_t* elems,
                                         int32_t* new_elems, int32_t* new_procs, pp_stream stream) {
  PP_REQUIRE(mesh && ps && ((elems && new_elems && new_procs) || ps->capacity == 0), "null argument");
  if (ps->capacity == 0) return PP_OK;
  k_set_unsafe<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), elems, mesh->aux, mesh->self_rank, new_elems, new_procs);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_gyro_scatter(pp_mesh* mesh, pp_ps* ps, const int32_t* v2v, double rmax,
                                     int32_t nrings, int32_t points_per_ring, double* scatter_w,
                                     pp_stream stream_) {
  PP_REQUIRE(mesh && ps && v2v && scatter_w, "null argument");
  PP_REQUIRE(mesh->dim == 2, "gyro scatter needs a 2D (triangle) mesh");
  PP_REQUIRE(nrings >= 2 && points_per_ring >= 1, "need at least two rings and one point per ring");
  PP_REQUIRE(ps->nelems == mesh->nelems, "particle structure and mesh disagree on nelems");
  cudaStream_t s = (cudaStream_t)stream_;
  PP_TIME(s, "gyro scatter");
  const int ne = mesh->nelems, nv = mesh->nverts, gnr = nrings, gppr = points_per_ring;
  // ring selection of gyroScatter.hpp:184-192 (the particle radius is the constant 1.125*ringWidth)
  const double ringWidth = rmax / gnr;
  const double ptclRadius = ringWidth * 1.125;
  int ringDown = 0;
  for (int i = 2; i <= gnr; i++) ringDown += (ptclRadius >= ringWidth * i);
  const int ringUp = ringDown + 1;
  PP_REQUIRE(ringUp < gnr, "ring index out of range");
  double* ring_accum;
  PP_TRY(pp_dev_alloc(&ring_accum, (size_t)gnr * nv, s));
  PP_CUDA(cudaMemsetAsync(ring_accum, 0, sizeof(double) * (size_t)gnr * nv, s));
  PP_CUDA(cudaMemsetAsync(scatter_w, 0, sizeof(double) * (size_t)nv, s));
  int* cnt = nullptr;
  const int kind = ps->cfg.kind;
  if (kind == PP_PS_DPS) {
    PP_TRY(pp_dev_alloc(&cnt, ne, s));
    PP_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ne, s));
    if (ps->capacity > 0)
      k_count_flat<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(ps->view(), cnt);
  }
  if (ps->nptcls > 0)
    k_rings<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(
        mesh->elem2verts, ne, cnt, ps->row_ppe, ps->element_to_row,
        kind == PP_PS_CSR ? ps->offsets : nullptr, gnr, ringDown, ringUp, ring_accum);
  const long npts = (long)nv * gnr * gppr;
  k_scatter_mapped<<<pp_div_up(npts, kBlock), kBlock, 0, s>>>(ring_accum, v2v, npts, gnr, gppr, scatter_w);
  PP_KERNEL_CHECK();
  pp_dev_free(ring_accum, s);
  pp_dev_free(cnt, s);
  return PP_OK;
}

extern "C" pp_status pp_gyro_ring_map(pp_mesh* mesh, double rmax, int32_t nrings,
                                      int32_t points_per_ring, double theta_deg, int32_t* map_out,
                                      pp_search_stats* stats_host, pp_stream stream_) {
  PP_REQUIRE(mesh && map_out, "null argument");
  PP_REQUIRE(mesh->dim == 2, "gyro ring mapping needs a 2D (triangle) mesh");
  PP_REQUIRE(nrings >= 1 && points_per_ring >= 1, "bad ring configuration");
  cudaStream_t s = (cudaStream_t)stream_;
  const long npts = (long)mesh->nverts * nrings * points_per_ring;
  PP_REQUIRE(npts < 0x7fffffffL, "too many ring points");
  double* tgt;
  int *start, *ids;
  uint32_t* mask;
  PP_TRY(pp_dev_alloc(&tgt, 3 * (size_t)npts, s));
  PP_TRY(pp_dev_alloc(&start, npts, s));
  PP_TRY(pp_dev_alloc(&ids, npts, s));
  PP_TRY(pp_dev_alloc(&mask, (npts + 31) / 32 + 1, s));
  k_ring_points<<<pp_div_up((npts + 31) / 32 * 32, kBlock), kBlock, 0, s>>>(
      mesh->coords, mesh->vert_first_elem, npts, nrings, points_per_ring, rmax, theta_deg, tgt, start, mask);
  PP_CUDA(cudaMemsetAsync(ids, 0xff, sizeof(int) * npts, s));   // elem_ids = -1: start from `start`
  // searchAndBuildMap (gyroScatter.hpp:25-90): search_mesh_2d with maxLoops = 100 on a throw-away
  // structure of nverts*nrings*ppr pseudo-particles; here the points are the slots of a flat view
  PsView v;
  v.kind = PP_PS_DPS; v.capacity = (int)npts; v.mask_bits = mask; v.slot_elem = start;
  v.offsets = v.slice_to_chunk = v.row_to_element = v.tile_slice = nullptr; v.C = 1; v.nslices = 0;
  pp_search_args a;
  a.variant = PP_SEARCH_2D_LEGACY; a.x_orig = nullptr; a.x_tgt = tgt; a.stride = npts;
  a.elem_ids = ids; a.elem_ids_empty = 0; a.require_intersection = 0; a.inter_faces = nullptr;
  a.inter_points = nullptr; a.looplimit = 100;
  PP_TRY(pp_search_view(mesh, v, &a, stats_host, s));
  k_ring_map<<<pp_div_up(npts, kBlock), kBlock, 0, s>>>(ids, mesh->elem2verts, npts, map_out);
  PP_KERNEL_CHECK();
  pp_dev_free(tgt, s); pp_dev_free(start, s); pp_dev_free(ids, s); pp_dev_free(mask, s);
  return PP_OK;
}

extern "C" pp_status pp_gyro_interleave(const double* fwd, const double* bkwd, int32_t nverts,
                                        double* sync_array, pp_stream stream) {
  PP_REQUIRE(fwd && bkwd && sync_array && nverts >= 0, "bad argument");
  if (nverts == 0) return PP_OK;
  k_interleave<<<pp_div_up(nverts, kBlock), kBlock, 0, (cudaStream_t)stream>>>(fwd, bkwd, nverts, sync_array);
  PP_KERNEL_CHECK();
  return PP_OK;
}
