// pp_internal.cuh -- shared internals of libpumipic_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "pumipic_b200.h"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void pp_set_error(const char* fmt, ...);
void pp_runtime_init();   // once per device: keep the async allocation pool cached
void* pp_pinned_scratch();  // 256 bytes of pinned host memory per thread (or null)

#define PP_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      pp_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                     \
                   cudaGetErrorString(e__));                                              \
      return PP_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define PP_REQUIRE(cond, msg)                                                             \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      pp_set_error("%s:%d: %s", __FILE__, __LINE__, msg);                                 \
      return PP_ERR_INVALID;                                                              \
    }                                                                                     \
  } while (0)

#define PP_TRY(call)                                                                      \
  do {                                                                                    \
    pp_status s__ = (call);                                                               \
    if (s__ != PP_OK) return s__;                                                         \
  } while (0)

#define PP_KERNEL_CHECK() PP_CUDA(cudaGetLastError())

static inline int pp_div_up(long a, long b) { return (int)((a + b - 1) / b); }

// Stream-ordered device allocation helpers.
template <class T>
static inline pp_status pp_dev_alloc(T** p, size_t n, cudaStream_t s) {
  *p = nullptr;
  if (n == 0) n = 1;
  PP_CUDA(cudaMallocAsync((void**)p, n * sizeof(T), s));
  return PP_OK;
}
template <class T>
static inline void pp_dev_free(T* p, cudaStream_t s) {
  if (p) cudaFreeAsync((void*)p, s);
}
// copy `n` elements from `src` (host or device per memspace) into fresh device memory
template <class T>
static inline pp_status pp_dev_import(T** dst, const T* src, size_t n, int memspace,
                                      cudaStream_t s) {
  PP_TRY(pp_dev_alloc(dst, n, s));
  if (n && src)
    PP_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(T),
                            memspace == PP_HOST ? cudaMemcpyHostToDevice
                                                : cudaMemcpyDeviceToDevice, s));
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// phase timers / NVTX ranges (csrc/pp_timing.cu).  A scope is always an NVTX range; with
// pp_timing_enable(1) it is also bracketed by two CUDA events on `s` whose elapsed time is
// accumulated under `label` (support/ppTiming.cpp RecordTime).
// ------------------------------------------------------------------------------------------
class PPTimeScope {
 public:
  PPTimeScope(cudaStream_t s, const char* label);
  ~PPTimeScope();
  PPTimeScope(const PPTimeScope&) = delete;
  PPTimeScope& operator=(const PPTimeScope&) = delete;
 private:
  cudaStream_t s_;
  cudaEvent_t a_;
  int index_;
};
const char* pp_kind_name(int kind);
#define PP_TIME_CAT2(a, b) a##b
#define PP_TIME_CAT(a, b) PP_TIME_CAT2(a, b)
#define PP_TIME(stream, label) PPTimeScope PP_TIME_CAT(pp_time_scope_, __LINE__)((stream), (label))
#define PP_TIME_KIND(stream, kind, what)                                                       \
  const std::string PP_TIME_CAT(pp_time_label_, __LINE__) = std::string(pp_kind_name(kind)) + " " + (what); \
  PPTimeScope PP_TIME_CAT(pp_time_scope_, __LINE__)((stream), PP_TIME_CAT(pp_time_label_, __LINE__).c_str())

// ------------------------------------------------------------------------------------------
// walk records: everything one hop of the adjacency walk needs, in one aligned gather.
// ------------------------------------------------------------------------------------------
// adj[f] >= 0 : element across local side f
// adj[f] <  0 : side is exposed (domain boundary); side id = -adj[f]-1
// codes: 8 bits per local side f:
//    3D: bits0-1,2-3,4-5 = tet-local index of the side's own vertices fv0,fv1,fv2
//        bit6 = isFaceFlipped (pumipic_utils.hpp:501-507), bit7 = legacy flip (adjacency.hpp:662-664)
//    2D: bits0-1,2-3 = triangle-local index of ev0,ev1; bit6 = isFaceFlipped (utils.hpp:495-499)
// aux: owner rank of the element if it is NOT safe on this PICpart, else -1 (pp_mesh_set_picpart)
struct __align__(16) PPTetRec {   // 128 B = one L2 line
  double c[12];                   // 4 vertices x 3 coordinates
  double vol;                     // measure_elements_real
  int adj[4];
  unsigned codes;
  int aux;
};
struct __align__(16) PPTriRec {   // 96 B = three 32 B sectors
  double c[6];                    // 3 vertices x 2 coordinates
  double area;
  int adj[3];
  unsigned codes;
  int cls;                        // element class id (ellipticalPush)
  int aux;
  int pad[2];
};
// BCC walk record of a tet: the particle-independent half of barycentric_tet
// (adjacency.tpp:41-69) computed once per element with the reference's own operations, so the
// per-particle part is 9 subtractions + 4 dot products + 4 multiplies.  12 x 16 B pieces.
struct __align__(16) PPBccRec3 {  // 192 B
  double a[9];                    // anchors of faces 0/1, 2, 3: M0, M1, M2
  double n[12];                   // n[f] = cross(c - a, b - a) of face f = (a, b, c)
  double inv_vol;                 // 1.0 / vol if vol > 0, else -1.0 (barycentric_tet fails)
  int adj[4];
};
static_assert(sizeof(PPBccRec3) == 192, "BCC walk record must be 192 bytes");
static_assert(sizeof(PPTetRec) == 128, "tet walk record must be 128 bytes");
static_assert(sizeof(PPTriRec) == 96, "tri walk record must be 96 bytes");

struct pp_mesh {
  int dim, nverts, nelems, nsides;
  double tol, min_measure;
  int n_exposed;
  int self_rank;
  // device arrays (owned)
  double* coords;
  int* elem2verts;
  int* elem2sides;
  int* side2verts;
  int* elem_class;
  double* measure;
  int8_t* exposed;
  int* side2elem;   // [2*nsides] (lo, hi|-1)
  int* dual_off;
  int* dual;
  int* safe;        // [nelems] or null
  int* owner;       // [nelems] or null
  void* walk;       // PPTetRec[nelems] or PPTriRec[nelems]
  PPBccRec3* walk_bcc;  // [nelems] (3D only)
  int* aux;         // [nelems] owner rank if the element is not safe here, else -1
  int* vert_first_elem;  // [nverts] lowest-numbered element adjacent to each vertex (ask_up(0,dim) first entry)
  // search scratch
  int* stats_dev;   // device counters (see SearchCounters)
  void* hostpipe;   // staging buffers / streams of pp_push_direction_search_host (lazy)
};

// device-side counters of one search
struct SearchCounters {
  int max_iters;
  int not_in_elem;
  int not_found;
  int aborted;
  int active;
  int next_chunk;   // dynamic chunk scheduler of the chunk walk (k_walk_scs)
  unsigned long long hops;
};

// ------------------------------------------------------------------------------------------
// particle structure
// ------------------------------------------------------------------------------------------
struct PsView {  // what kernels need to map slot -> (row element, mask)
  int kind;
  int capacity;
  const uint32_t* mask_bits;
  const int* slot_elem;  // DPS / CSR (and materialised SCS): element per slot, or null
  // SCS geometry
  const int* offsets;
  const int* slice_to_chunk;
  const int* row_to_element;
  const int* tile_slice;  // slice holding the first slot of each 32-slot tile
  int C;
  int nslices;
  // chunk geometry (SCS): slots of chunk c are [chunk_start[c], chunk_start[c+1]), column-major
  // over its C rows: slot = chunk_start[c] + col*C + row
  const int* chunk_start = nullptr;   // [nchunks+1]
  int nchunks = 0;
  int nelems = 0;
  int first_chunk = 0;                // chunks [0, first_chunk) hold no particle
  int sliced = 0;                     // some chunk is wider than V columns: it spans several slices
};

struct pp_ps {
  pp_ps_config cfg;
  int nmembers;
  std::vector<pp_member_desc> members;
  int nelems, nptcls, capacity, nrows;
  long stride;              // allocated slots per member component (>= capacity)
  std::vector<void*> data;  // one device array per member: [ncomp][stride]
  std::vector<void*> swap;  // SCS double buffer
  long swap_stride;
  long data_alloc = 0, swap_alloc = 0;   // slots per component the SCS arrays were allocated with (>= stride)
  uint32_t* mask_bits;
  long mask_words_alloc;
  int* slot_elem;           // [capacity] (DPS parent array; CSR/SCS materialised map)
  bool slot_elem_valid;     // kernels read slot_elem (DPS / CSR)
  bool slot_elem_materialized;  // SCS: slot_elem filled on demand for pp_ps_get_layout
  // SCS
  int C, V, nchunks, nslices;
  int* offsets;
  int* slice_to_chunk;
  int* row_to_element;
  int* element_to_row;
  int* tile_slice;
  int* chunk_start;         // [nchunks+1] first slot of each chunk, chunk_start[nchunks] = capacity
  int* row_ppe;             // [nrows] particles per row at the last (re)build
  int64_t* elem_gids;       // [nelems] or null
  int64_t* sorted_gid;      // [nelems] gids ascending (built lazily for migrate)
  int* sorted_lid;          // [nelems] local id of sorted_gid[i]
  char* stage;              // record stage of the rebuild (grow-only scratch)
  int shuffle_skip = 0;     // rebuilds to wait before the next reshuffle attempt
  int shuffle_streak = 0;   // consecutive failed reshuffle attempts
  std::vector<int> rebuild_remap;   // one-shot member remap of the next rebuild (pp_ps_set_rebuild_remap)
  int first_chunk = 0;      // chunks before this one are empty (single sort window: empty rows lead)
  int sliced = 0;           // some chunk spans several slices (more slices than non-empty chunks)
  int ppe_bits_hint = 0;    // key bits of the row sort guessed from the last rebuild's largest row (0 = none)
  int nnz_hint = 0;         // non-empty rows of the last rebuild (0 = none): a mostly empty structure only sorts those
  size_t stage_bytes;
  PsView view() const;
};

// pp_ps_rebuild with a device-side count of the particles being added (see csrc/pp_scs.cu)
pp_status pp_ps_rebuild_ex(pp_ps* ps, const int32_t* new_element, int32_t n_new, const int32_t* n_new_dev,
                           int64_t new_ld, const int32_t* new_particle_elements,
                           const void* const* new_particle_info, pp_stream stream);

// slot -> (element of owning row, mask).  Works for every structure kind.
__device__ __forceinline__ bool pp_slot_lookup(const PsView& v, int slot, int& elem) {
  const uint32_t w = __ldg(v.mask_bits + (slot >> 5));
  const bool m = (w >> (slot & 31)) & 1u;
  if (v.slot_elem) {
    elem = __ldg(v.slot_elem + slot);
  } else {
    int S = __ldg(v.tile_slice + (slot >> 5));
    while (slot >= __ldg(v.offsets + S + 1)) ++S;
    const int r = (slot - __ldg(v.offsets + S)) % v.C;
    elem = __ldg(v.row_to_element + __ldg(v.slice_to_chunk + S) * v.C + r);
  }
  return m;
}

// ------------------------------------------------------------------------------------------
// Omega_h small-vector arithmetic, same evaluation order as the reference's CPU path.
// The library is compiled with -fmad=false so no multiply-add is ever contracted.
// ------------------------------------------------------------------------------------------
struct d3 { double x, y, z; };
struct d2 { double x, y; };
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot3(d3 a, d3 b) {
  double c = a.x * b.x;
  c = c + a.y * b.y;
  c = c + a.z * b.z;
  return c;
}
__device__ __forceinline__ double norm3(d3 a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ d2 operator-(d2 a, d2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ double dot2(d2 a, d2 b) {
  double c = a.x * b.x;
  c = c + a.y * b.y;
  return c;
}
__device__ __forceinline__ double cross2(d2 a, d2 b) { return a.x * b.y - a.y * b.x; }

// Omega_h::are_close(a, 0, tol, tol) || a > 0   (pumipic_utils.hpp:78-86)
// are_close(a,0,tol,tol): |a| <= tol, else |0-a|/max(|a|,0) = |a|/|a| <= tol, which is 1 <= tol for
// finite a and NaN (false) for infinite a -- evaluated without the division.
__device__ __forceinline__ bool pp_gtez(double a, double tol) {
#ifndef PP_AB_GTEZ_V1
  // for 0 <= tol < 1 (every tolerance of the path: 1e-20, 1e-10, max(1e-15/min_area, 1e-8)) the
  // test below is (|a| <= tol) || (a > 0), i.e. exactly a >= -tol for every a including NaN and
  // the infinities: one compare instead of three
  if (tol >= 0.0 && tol < 1.0) return a >= -tol;
#endif
  const double am = fabs(a);
#ifdef PP_AB_OLD_GTEZ
  bool close;
  if (am <= tol) close = true;
  else close = (am / am) <= tol;
#else
  const bool close = (am <= tol) || (tol >= 1.0 && am < CUDART_INF);
#endif
  return close || a > 0;
}
