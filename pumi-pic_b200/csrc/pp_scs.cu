// pp_scs.cu -- Sell-C-sigma layout construction and the rebuild of every structure kind.
//
// Replaces particle_structs/src/scs: chooseChunkHeight SCS_buildFns.h:4-16, sigmaSort
// SCS_sort.h:4-49 (CUDA branch: ascending thrust::sort_by_key per sigma window),
// constructChunks :19-98 (with the three padding strategies), constructOffsets :115-153,
// setupParticleMask :155-199, initSCSData :202-225 and rebuild SCS_rebuild.h:123-314;
// CSR_rebuild.hpp:18-118 and dps_rebuild.hpp for the flat kinds.
//
// B200-first differences: (1) the mask is a bit per slot; (2) a rebuild moves ALL members of a
// particle in one kernel (the reference launches one gather/scatter kernel per member type);
// (3) the per-32-slot `tile_slice` table lets any kernel map slot -> row without the
// slice-per-team launch shape of SellCSigma::parallel_for.
#include <algorithm>

#include <cub/cub.cuh>

#include "pp_internal.cuh"

pp_status pp_ps_alloc_members(pp_ps* ps, std::vector<void*>& arrs, long stride, cudaStream_t s);

namespace {
constexpr int kBlock = 256;

struct ScsLayout {
  int C = 1, nchunks = 0, nrows = 0, nslices = 0, capacity = 0;
  int* offsets = nullptr;
  int* slice_to_chunk = nullptr;
  int* row_to_element = nullptr;
  int* element_to_row = nullptr;
  int* chunk_start = nullptr;   // first slot of each chunk
  int* tile_slice = nullptr;
  int* row_ppe = nullptr;       // particles per row (sorted order)
  uint32_t* mask = nullptr;
  long mask_words = 0;
  int nonempty_chunks = -1;     // chunks that hold particles (-1: unknown)
};

// grid-stride, one atomic per block: per-warp atomics on one address serialise (25 us for 1 M elements)
__global__ void k_count_nonzero(const int* __restrict__ a, int n, int* out) {
  __shared__ int part[kBlock / 32];
  int cnt = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    cnt += a[i] > 0;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kBlock / 32; ++w) t += part[w];
    if (t) atomicAdd(out, t);
  }
}

// sigmaSort keys: ascending particle count inside windows of `sigma` elements (stable)
template <class Key>
__global__ void k_sort_keys(const int* __restrict__ ppe, int ne, int sigma, int cbits, Key* keys, int* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ne) return;
  const Key win = (Key)(i / sigma);
  keys[i] = (win << cbits) | (Key)(uint32_t)ppe[i];
  vals[i] = i;
}

__global__ void k_rows(const int* __restrict__ sorted_elem, const int* __restrict__ ppe, int ne,
                       int nrows, int* row2elem, int* elem2row, int* row_ppe) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  if (i < ne) {
    const int e = sorted_elem ? sorted_elem[i] : i;
    row2elem[i] = e;
    elem2row[e] = i;
    row_ppe[i] = ppe[e];
  } else {               // padding rows up to a multiple of C (SCS_buildFns.h:39-44)
    row2elem[i] = i;
    elem2row[i] = i;
    row_ppe[i] = 0;
  }
}

// chunk width = widest row; cw[0]=sum, cw[1]=count of non-empty chunks; inv = sum of 1/width.
// One warp per chunk (lanes stride over the C rows), warps stride over the chunks and keep their
// partial sums in registers: three atomics per block instead of three per chunk.
__global__ void k_chunk_widths(const int* __restrict__ row_ppe, int nchunks, int C, int* width,
                               int* cw, double* inv) {
  __shared__ int p_sum[kBlock / 32], p_cnt[kBlock / 32];
  __shared__ double p_inv[kBlock / 32];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  int sum = 0, cnt = 0;
  double isum = 0.0;
  for (long c = blockIdx.x * (long)(kBlock / 32) + wid; c < nchunks; c += (long)gridDim.x * (kBlock / 32)) {
    int w = 0;
    for (int r = lane; r < C; r += 32) w = max(w, row_ppe[c * C + r]);
    w = __reduce_max_sync(0xffffffffu, w);
    if (lane == 0) {
      width[c] = w;
      if (w > 0) { sum += w; cnt += 1; isum += 1.0 / w; }
    }
  }
  if (lane == 0) { p_sum[wid] = sum; p_cnt[wid] = cnt; p_inv[wid] = isum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { sum += p_sum[w]; cnt += p_cnt[w]; isum += p_inv[w]; }
    if (cnt) {
      atomicAdd(cw, sum);
      atomicAdd(cw + 1, cnt);
      atomicAdd(inv, isum);
    }
  }
}

// SCS_buildFns.h:62-97
__global__ void k_pad_widths(int* width, int nchunks, const int* __restrict__ cw,
                             const double* __restrict__ inv, double pad, int strat) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const int cw_sum = cw[0], cw_cnt = cw[1];
  if (cw_sum <= 0) return;
  const int w = width[c];
  if (strat == PP_PAD_EVENLY) {
    const int avg_pad = (int)(cw_sum * pad / cw_cnt);
    if (w > 0) width[c] = w + avg_pad;
  } else if (strat == PP_PAD_PROPORTIONALLY) {
    width[c] = (int)(w + w * pad);
  } else {
    const double cw_sum2 = cw_sum / inv[0] * pad;
    if (w != 0) width[c] = (int)(w + cw_sum2 / w);
  }
}

__global__ void k_slices_per_chunk(const int* __restrict__ width, int nchunks, int V, int* spc) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nchunks) return;
  spc[c] = c < nchunks ? width[c] / V + (width[c] % V != 0) : 0;
}

__global__ void k_fill_slices(const int* __restrict__ width, const int* __restrict__ slice_off,
                              int nchunks, int V, int C, int* s2c, int* slice_size) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const int b = slice_off[c], e = slice_off[c + 1];
  for (int j = b; j < e; ++j) {
    s2c[j] = c;
    const int rem = width[c] % V;
    const int last = rem + (rem == 0) * V;
    slice_size[j] = (j == e - 1) ? last * C : V * C;
  }
}

__global__ void k_chunk_start(const int* __restrict__ slice_off, const int* __restrict__ offsets,
                              int nchunks, int capacity, int* chunk_start) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nchunks) return;
  // an empty chunk owns no slice: its (empty) slot range sits where the next chunk begins
  chunk_start[c] = c < nchunks ? offsets[slice_off[c]] : capacity;
}

__global__ void k_tile_slice(const int* __restrict__ offsets, int nslices, int ntiles, int* tile_slice) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int slot = t * 32;
  int lo = 0, hi = nslices;   // last S with offsets[S] <= slot
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= slot) lo = mid; else hi = mid;
  }
  tile_slice[t] = lo;
}

// setupParticleMask (SCS_buildFns.h:155-199): slot (row, col) holds a particle iff col < ppe(row)
__global__ void k_scs_mask(PsView v, const int* __restrict__ row_ppe, int ne, uint32_t* mask, long nwords) {
  const long s = blockIdx.x * (long)blockDim.x + threadIdx.x;
  bool bit = false;
  if (s < v.capacity) {
    int S = v.tile_slice[s >> 5];
    while (s >= v.offsets[S + 1]) ++S;
    const int rel = (int)(s - v.offsets[S]);
    const int r = rel % v.C;
    // column inside the chunk = columns of earlier slices of the same chunk + column in slice
    const int chunk = v.slice_to_chunk[S];
    int S0 = S;
    while (S0 > 0 && v.slice_to_chunk[S0 - 1] == chunk) --S0;
    const int col = (int)((s - v.offsets[S0]) / v.C);
    const int row = chunk * v.C + r;
    bit = v.row_to_element[row] < ne && col < row_ppe[row];
  }
  const unsigned m = __ballot_sync(0xffffffffu, bit);
  if ((threadIdx.x & 31) == 0 && (s >> 5) < nwords) mask[s >> 5] = m;
}

__global__ void k_assign_slots(const int* __restrict__ elems, int n, const int* __restrict__ elem2row,
                               const int* __restrict__ chunk_start, int C, int* row_fill, int* slots,
                               const int* __restrict__ n_dev = nullptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  const int row = elem2row[elems[i]];
  const int col = atomicAdd(row_fill + row, 1);
  slots[i] = chunk_start[row / C] + row % C + col * C;
}

// ---- rebuild kernels
// countNewParticles (SCS_rebuild.h:133-138).  The value the atomic returns is the particle's rank
// inside its destination element, which is all the slot claim of the record move needs: with
// `rank` the move runs without a second round of atomics.
__global__ void k_hist_kept(PsView v, const int* __restrict__ new_elem, int* count, int* rank) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int e = -1;
  if (s < v.capacity) {
    const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
    if (m) e = new_elem[s];
  }
  if (e >= 0) {
    if (rank) rank[s] = atomicAdd(count + e, 1);
    else atomicAdd(count + e, 1);
  }
}
// The same with the destinations of a block's kHistSlots slots counted in shared memory first: a block
// reserves the ranks of all its particles that go to one element with ONE global atomic (its return
// value + the particle's rank inside the block = the particle's rank in the element).  Particles of a
// row mostly stay in their element, so the 32 rows a block sees fold their stayers into 32 atomics; a
// row that holds a large share of all particles (pseudoXGCm's load: ~1 M in one element) no longer
// serialises a million atomics on one counter.
constexpr int kHistPer = 4, kHistSlots = 256 * kHistPer, kHistTable = 2048, kHistProbes = 48;
__global__ void __launch_bounds__(256) k_hist_kept_block(PsView v, const int* __restrict__ new_elem, int* count,
                                                         int* rank) {
  __shared__ int s_key[kHistTable], s_cnt[kHistTable], s_base[kHistTable];
  for (int i = threadIdx.x; i < kHistTable; i += 256) { s_key[i] = -1; s_cnt[i] = 0; }
  __syncthreads();
  int el[kHistPer], ent[kHistPer], pos[kHistPer];
  const long s0 = (long)blockIdx.x * kHistSlots + threadIdx.x;
#pragma unroll
  for (int j = 0; j < kHistPer; ++j) {
    const long s = s0 + (long)j * 256;
    int e = -1;
    if (s < v.capacity) {
      const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
      if (m) e = new_elem[s];
    }
    el[j] = e; ent[j] = -1; pos[j] = 0;
    if (e >= 0) {
      unsigned h = ((unsigned)e * 2654435761u) >> 21;            // 11 bits
      for (int probe = 0; probe < kHistProbes; ++probe) {
        const int k = atomicCAS(&s_key[h], -1, e);
        if (k == -1 || k == e) { ent[j] = (int)h; break; }
        h = (h + 1) & (kHistTable - 1);
      }
      if (ent[j] >= 0) pos[j] = atomicAdd(&s_cnt[ent[j]], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHistTable; i += 256)
    if (s_key[i] >= 0) s_base[i] = atomicAdd(count + s_key[i], s_cnt[i]);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kHistPer; ++j) {
    if (el[j] < 0) continue;
    const long s = s0 + (long)j * 256;
    const int r = ent[j] >= 0 ? s_base[ent[j]] + pos[j] : atomicAdd(count + el[j], 1);   // table full: on its own
    if (rank) rank[s] = r;
  }
}
// n_dev (all *_new kernels): the number of particles being added when only the device knows it
// (pp_ps_migrate over the peer-memory window); n is then an upper bound that sized the launch
// rank_out: the particle's rank in its element, behind the kept particles counted before (stream order)
__global__ void k_hist_new(const int* __restrict__ elems, int n, int* count, int* bad,
                           const int* __restrict__ n_dev = nullptr, int* rank_out = nullptr) {
  if (n_dev) n = min(n, *n_dev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = elems[i];
    if (e < 0) { *bad = 1; if (rank_out) rank_out[i] = 0; continue; }
    const int r = atomicAdd(count + e, 1);
    if (rank_out) rank_out[i] = r;
  }
}
// new particles: src_of[slot of (element, rank)] = -(i + 1)
__global__ void k_invmap_new_ranked(const int* __restrict__ elems, const int* __restrict__ rank, int n,
                                    const int* __restrict__ elem2row, const int* __restrict__ chunk_start,
                                    int* src_of, int* slots, const int* __restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = elems[i];
    if (e < 0) continue;
    const int row = __ldg(elem2row + e);
    const int slot = __ldg(chunk_start + (row >> 5)) + (rank[i] << 5) + (row & 31);
    if (src_of) src_of[slot] = -i - 1;
    if (slots) slots[i] = slot;
  }
}

struct MemberTable {
  int n;
  char* src[16];
  char* dst[16];
  int bytes[16];   // scalar bytes
  int ncomp[16];
};

// one kernel moves every member of every kept particle (CopyPSToPS psMemberType.h:74-115 moves
// one member per kernel launch)
__global__ void k_move_kept(PsView v, const int* __restrict__ new_elem,
                            const int* __restrict__ elem2row, const int* __restrict__ chunk_start,
                            int C, int dense, const int* __restrict__ dense_off, int* row_fill,
                            MemberTable mt, long src_stride, long dst_stride, uint32_t* new_mask_unused,
                            int* new_slot_of) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  int ns = -1;
  if (m) {
    const int e = new_elem[s];
    if (e >= 0) {
      if (dense) {
        ns = dense_off[e] + atomicAdd(row_fill + e, 1);
      } else {
        const int row = elem2row[e];
        const int col = atomicAdd(row_fill + row, 1);
        ns = chunk_start[row / C] + row % C + col * C;
      }
      for (int k = 0; k < mt.n; ++k) {
        const int sb = mt.bytes[k];
        for (int c = 0; c < mt.ncomp[k]; ++c) {
          const char* a = mt.src[k] + ((long)c * src_stride + s) * sb;
          char* b = mt.dst[k] + ((long)c * dst_stride + ns) * sb;
          if (sb == 8) *(double*)b = *(const double*)a;
          else if (sb == 4) *(int*)b = *(const int*)a;
          else for (int q = 0; q < sb; ++q) b[q] = a[q];
        }
      }
    }
  }
  if (new_slot_of) new_slot_of[s] = ns;
}

__global__ void k_place_new(const int* __restrict__ slots, int n, MemberTable mt, long dst_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ns = slots[i];
  for (int k = 0; k < mt.n; ++k) {
    const int sb = mt.bytes[k];
    for (int c = 0; c < mt.ncomp[k]; ++c) {
      const char* a = mt.src[k] + ((long)c * n + i) * sb;   // new particle arrays are [ncomp][n]
      char* b = mt.dst[k] + ((long)c * dst_stride + ns) * sb;
      if (sb == 8) *(double*)b = *(const double*)a;
      else if (sb == 4) *(int*)b = *(const int*)a;
      else for (int q = 0; q < sb; ++q) b[q] = a[q];
    }
  }
}

__global__ void k_assign_dense(const int* __restrict__ elems, int n, const int* __restrict__ off,
                               int* fill, int* slots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = elems[i];
  slots[i] = off[e] + atomicAdd(fill + e, 1);
}

__global__ void k_mask_first_n2(uint32_t* mask, long nwords, int n) {
  long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  long lo = w * 32;
  uint32_t v = 0;
  if (lo + 32 <= n) v = 0xffffffffu;
  else if (lo < n) v = (1u << (n - lo)) - 1u;
  mask[w] = v;
}

__global__ void k_expand_offsets2(const int* __restrict__ off, int ne, int n, int cap, int* slot_elem) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  if (s >= n) { slot_elem[s] = 0; return; }
  int lo = 0, hi = ne;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= s) lo = mid; else hi = mid;
  }
  slot_elem[s] = lo;
}

// DPS rebuild: particles stay where they are; deleted slots become holes that new particles fill
__global__ void k_dps_update(PsView v, const int* __restrict__ new_elem, int* slot_elem,
                             uint32_t* mask, int* nkept) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (s < v.capacity) {
    const bool m = (mask[s >> 5] >> (s & 31)) & 1u;
    if (m) {
      const int e = new_elem[s];
      if (e >= 0) { keep = true; slot_elem[s] = e; }
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) {
    if ((s >> 5) < (v.capacity + 31) / 32) mask[s >> 5] = b;
    if (b) atomicAdd(nkept, __popc(b));
  }
}
// rank the free slots: hole h (0-based among unset mask bits in slot order) receives new particle h
__global__ void k_dps_hole_count(const uint32_t* __restrict__ mask, long nwords, int cap, int* cnt) {
  long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (w > nwords) return;
  if (w == nwords) { cnt[w] = 0; return; }
  uint32_t free_bits = ~mask[w];
  const long lo = w * 32;
  if (lo + 32 > cap) free_bits &= (cap > lo) ? ((1u << (cap - lo)) - 1u) : 0u;
  cnt[w] = __popc(free_bits);
}
__global__ void k_dps_fill_holes(uint32_t* mask, long nwords, int cap, const int* __restrict__ hole_off,
                                 const int* __restrict__ new_elems, int n_new, int* slot_elem, int* slots) {
  long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t free_bits = ~mask[w];
  const long lo = w * 32;
  if (lo + 32 > cap) free_bits &= (cap > lo) ? ((1u << (cap - lo)) - 1u) : 0u;
  int h = hole_off[w];
  uint32_t set = 0;
  while (free_bits && h < n_new) {
    const int b = __ffs(free_bits) - 1;
    free_bits &= free_bits - 1;
    const int s = (int)(lo + b);
    slots[h] = s;
    slot_elem[s] = new_elems[h];
    set |= 1u << b;
    ++h;
  }
  if (set) mask[w] |= set;
}

// ------------------------------------------------------------------------------------------
// Staged record move (the full re-layout of an element-sorted structure).
//
// With rows re-sorted by particle count, the destination slot of a particle is unrelated to its
// source slot, and in the component-major SoA layout every 8-byte component of a particle sits
// in a different 32-byte sector: a direct scatter turns each component store into a partial
// sector write at a random address.  The move therefore goes through an array-of-records stage:
//   pack   : read the old structure in slot order (coalesced), claim the destination slot, write
//            the whole record (padded to full sectors) at stage[dest];
//   unpack : read the stage in destination order and write the new SoA columns coalesced; the
//            particle mask of the new structure is produced by the same pass.
// Every DRAM access is a full sector; traffic is 2 x record (SoA) + 2 x padded record (stage).
// ------------------------------------------------------------------------------------------
constexpr int kMaxUnits = 40;          // 8-byte units of a particle record (<= 320 B)
struct UnitTable {
  int nunits;                          // units in use
  int rec_bytes;                       // bytes of a record that are moved (multiple of 32)
  int rec_stride;                      // record stride in the stage (rec_bytes, or padded to 128-byte lines)
  unsigned char kind[kMaxUnits];       // 0: one 8-byte scalar, 1: two 4-byte scalars (b may be null)
  const char* sa[kMaxUnits];           // source component bases (slot 0)
  const char* sb[kMaxUnits];
  char* da[kMaxUnits];                 // destination component bases (slot 0)
  char* db[kMaxUnits];
};

// a null source is a member that the rebuild zero-fills (pp_ps_set_rebuild_remap)
__device__ __forceinline__ unsigned long long unit_load(const UnitTable& t, int u, long s) {
  if (t.kind[u] == 0) return t.sa[u] ? *reinterpret_cast<const unsigned long long*>(t.sa[u] + 8 * s) : 0ull;
  const unsigned lo = t.sa[u] ? *reinterpret_cast<const unsigned*>(t.sa[u] + 4 * s) : 0u;
  const unsigned hi = t.sb[u] ? *reinterpret_cast<const unsigned*>(t.sb[u] + 4 * s) : 0u;
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
__device__ __forceinline__ void unit_store(const UnitTable& t, int u, long s, unsigned long long v) {
  if (t.kind[u] == 0) { *reinterpret_cast<unsigned long long*>(t.da[u] + 8 * s) = v; return; }
  *reinterpret_cast<unsigned*>(t.da[u] + 4 * s) = (unsigned)v;
  if (t.db[u]) *reinterpret_cast<unsigned*>(t.db[u] + 4 * s) = (unsigned)(v >> 32);
}

constexpr int kGroup = 4;   // quads (16 B) moved per batch: 8 independent loads in flight per thread

__device__ __forceinline__ void load_group(const UnitTable& t, int q0, long s, uint4 (&r)[kGroup]) {
#pragma unroll
  for (int k = 0; k < kGroup; ++k) {
    const int u = 2 * (q0 + k);
    const unsigned long long a = u < t.nunits ? unit_load(t, u, s) : 0ull;
    const unsigned long long b = u + 1 < t.nunits ? unit_load(t, u + 1, s) : 0ull;
    r[k] = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
  }
}
__device__ __forceinline__ void pack_record(const UnitTable& t, long s, uint4* rec, const uint4 (&first)[kGroup]) {
  const int nq = t.rec_bytes >> 4;
#pragma unroll
  for (int k = 0; k < kGroup; ++k)
    if (k < nq) rec[k] = first[k];
  for (int q0 = kGroup; q0 < nq; q0 += kGroup) {
    uint4 r[kGroup];
    load_group(t, q0, s, r);
#pragma unroll
    for (int k = 0; k < kGroup; ++k)
      if (q0 + k < nq) rec[q0 + k] = r[k];
  }
}

// pack kept particles: dense != 0 -> CSR destination (dense_off[e] + fill), else Sell-C-sigma.
//   * the first batch of record loads is issued before the destination chain (new element -> row
//     -> slot claim), so the two latencies overlap;
//   * records are written cooperatively: every lane parks its record in shared memory, then
//     Q = rec_bytes/16 consecutive lanes store one record as one contiguous run.  A per-lane
//     store of 16 bytes to 32 random records costs 32 memory transactions per instruction and the
//     kernel becomes transaction-bound (measured: lg_throttle, 18 % of DRAM peak at 160 B records).
// warp-level body: lane packs slot s (m = slot holds a particle), then the warp stores the records
__device__ __forceinline__ void stage_pack_warp(int s, bool m, const int* __restrict__ new_elem,
                                                const int* __restrict__ elem2row,
                                                const int* __restrict__ chunk_start, int C, int dense,
                                                const int* __restrict__ dense_off, int* row_fill,
                                                const int* __restrict__ rank, const UnitTable& t, char* stage,
                                                unsigned char* wsm, int lane) {
  const int stride = t.rec_bytes + 16;                       // bank-conflict-free row stride
  const int nq = t.rec_bytes >> 4;
  int ns = -1;
  if (m) {
    const int e = __ldg(new_elem + s);
    uint4 r[kGroup];
    load_group(t, 0, s, r);
    if (e >= 0) {
      if (dense) {
        ns = __ldg(dense_off + e) + (rank ? __ldg(rank + s) : atomicAdd(row_fill + e, 1));
      } else {
        const int row = __ldg(elem2row + e);
        const int cs = __ldg(chunk_start + row / C);
        const int col = rank ? __ldg(rank + s) : atomicAdd(row_fill + row, 1);
        ns = cs + row % C + col * C;
      }
      uint4* mine = reinterpret_cast<uint4*>(wsm + lane * stride);
#pragma unroll
      for (int k = 0; k < kGroup; ++k)
        if (k < nq) mine[k] = r[k];
      for (int q0 = kGroup; q0 < nq; q0 += kGroup) {
        load_group(t, q0, s, r);
#pragma unroll
        for (int k = 0; k < kGroup; ++k)
          if (q0 + k < nq) mine[q0 + k] = r[k];
      }
    }
  }
  __syncwarp();
  const int rpi = 32 / nq;                                   // records per store instruction
  const int myrec = lane / nq, piece = lane - myrec * nq;
  const bool act = myrec < rpi;
  for (int base = 0; base < 32; base += rpi) {
    const int rec = base + myrec;
    const int nsr = __shfl_sync(0xffffffffu, ns, rec & 31);
    if (act && rec < 32 && nsr >= 0)
      *reinterpret_cast<uint4*>(stage + (long)nsr * t.rec_stride + piece * 16) =
          *reinterpret_cast<const uint4*>(wsm + rec * stride + piece * 16);
  }
  __syncwarp();
}
__global__ void __launch_bounds__(256) k_stage_pack(PsView v, const int* __restrict__ new_elem,
                                                    const int* __restrict__ elem2row,
                                                    const int* __restrict__ chunk_start, int C, int dense,
                                                    const int* __restrict__ dense_off, int* row_fill,
                                                    const int* __restrict__ rank,
                                                    const __grid_constant__ UnitTable t, char* stage) {
  extern __shared__ __align__(16) unsigned char pack_smem[];
  const int lane = threadIdx.x & 31;
  unsigned char* wsm = pack_smem + (threadIdx.x >> 5) * 32 * (t.rec_bytes + 16);
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool m = false;
  if (s < v.capacity) m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (__ballot_sync(0xffffffffu, m) == 0u) return;           // warp-uniform
  stage_pack_warp(s, m, new_elem, elem2row, chunk_start, C, dense, dense_off, row_fill, rank, t, stage, wsm, lane);
}
// pack new particles: member arrays are [ncomp][n], destination slots precomputed
__global__ void __launch_bounds__(256) k_stage_pack_new(const int* __restrict__ slots, int n,
                                                        const __grid_constant__ UnitTable t, char* stage,
                                                        const int* __restrict__ n_dev = nullptr) {
  if (n_dev) n = min(n, *n_dev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint4 first[kGroup];
    load_group(t, 0, i, first);
    pack_record(t, i, reinterpret_cast<uint4*>(stage + (long)slots[i] * t.rec_stride), first);
  }
}
__device__ __forceinline__ void unpack_record(const UnitTable& t, const char* stage, long slot) {
  const uint4* rec = reinterpret_cast<const uint4*>(stage + slot * t.rec_stride);
  const int nq = (t.nunits + 1) >> 1;
  for (int q0 = 0; q0 < nq; q0 += kGroup) {
    uint4 r[kGroup];
#pragma unroll
    for (int k = 0; k < kGroup; ++k)
      if (q0 + k < nq) r[k] = __ldg(rec + q0 + k);     // cached: neighbouring quads share sectors
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
      const int u = 2 * (q0 + k);
      if (u < t.nunits) unit_store(t, u, slot, (unsigned long long)r[k].x | ((unsigned long long)r[k].y << 32));
      if (u + 1 < t.nunits) unit_store(t, u + 1, slot, (unsigned long long)r[k].z | ((unsigned long long)r[k].w << 32));
    }
  }
}
// unpack into a Sell-C-sigma layout with C = 32: one warp per 32-slot tile (= one column of a
// chunk, lane = row); also writes the particle mask (slot (row, col) holds a particle iff
// col < ppe(row), SCS_buildFns.h:155-199)
__global__ void __launch_bounds__(256) k_stage_unpack_scs(PsView v, const int* __restrict__ row_ppe,
                                                          const __grid_constant__ UnitTable t,
                                                          const char* __restrict__ stage, uint32_t* mask) {
  const int lane = threadIdx.x & 31;
  const long tile = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long sb = tile * 32;
  if (sb >= v.capacity) return;
  int S = __ldg(v.tile_slice + tile);
  while (sb >= __ldg(v.offsets + S + 1)) ++S;
  const int chunk = __ldg(v.slice_to_chunk + S);
  const int col = (int)((sb - __ldg(v.chunk_start + chunk)) >> 5);
  const bool valid = col < __ldg(row_ppe + chunk * 32 + lane);
  const unsigned w = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) mask[tile] = w;
  if (valid) unpack_record(t, stage, sb + lane);
}
// unpack into a dense (CSR) layout: the first `n` slots hold particles
__global__ void __launch_bounds__(256) k_stage_unpack_dense(int n, const __grid_constant__ UnitTable t,
                                                            const char* __restrict__ stage) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) unpack_record(t, stage, s);
}

pp_status launch_stage_pack(const PsView& v, const int* new_elem, const int* elem2row, const int* chunk_start,
                            int C, int dense, const int* dense_off, int* row_fill, const int* rank,
                            const UnitTable& t, char* stage, cudaStream_t s) {
  const size_t smem = (size_t)(kBlock / 32) * 32 * (t.rec_bytes + 16);
  PP_CUDA(cudaFuncSetAttribute(k_stage_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_stage_pack<<<pp_div_up(v.capacity, kBlock), kBlock, smem, s>>>(v, new_elem, elem2row, chunk_start, C, dense,
                                                                   dense_off, row_fill, rank, t, stage);
  return PP_OK;
}

// Build the unit table of a structure; false if a member cannot be expressed in 4/8-byte units.
// remap (or null): destination member i is read from source member remap[i], -1 = zero-filled
bool unit_table(const pp_ps* ps, const void* const* src, long src_stride, const std::vector<void*>* dst,
                long dst_stride, UnitTable& t, const int* remap = nullptr) {
  t.nunits = 0;
  int n4 = 0;
  // 8-byte scalars first, then pairs of 4-byte scalars: natural alignment inside the record
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < ps->nmembers; ++i) {
      const int sb = ps->members[i].scalar_bytes;
      if (sb != 4 && sb != 8) return false;
      if ((pass == 0) != (sb == 8)) continue;
      for (int c = 0; c < ps->members[i].ncomp; ++c) {
        const int si = remap ? remap[i] : i;
        const char* sp = (src && si >= 0) ? (const char*)src[si] + (size_t)c * src_stride * sb : nullptr;
        char* dp = dst ? (char*)(*dst)[i] + (size_t)c * dst_stride * sb : nullptr;
        if (sb == 8) {
          if (t.nunits >= kMaxUnits) return false;
          t.kind[t.nunits] = 0; t.sa[t.nunits] = sp; t.sb[t.nunits] = nullptr;
          t.da[t.nunits] = dp; t.db[t.nunits] = nullptr; ++t.nunits;
        } else if (n4 % 2 == 0) {
          if (t.nunits >= kMaxUnits) return false;
          t.kind[t.nunits] = 1; t.sa[t.nunits] = sp; t.sb[t.nunits] = nullptr;
          t.da[t.nunits] = dp; t.db[t.nunits] = nullptr; ++t.nunits; ++n4;
        } else {
          t.sb[t.nunits - 1] = sp; t.db[t.nunits - 1] = dp; ++n4;
        }
      }
    }
  t.rec_bytes = ((t.nunits * 8 + 31) / 32) * 32;
  t.rec_stride = t.rec_bytes;
  return t.nunits > 0;
}

// ---- contention-free ranks for crowded elements.  With thousands of particles per element the
// per-element atomics of the histogram and of the slot claim serialise; instead the kept
// particles are sorted by destination element (stable radix sort on log2(ne) bits) and the rank
// of a particle inside its element, the per-element counts and the slot order all follow from
// the sorted sequence.  Side effect: the slot order inside a row is deterministic.
__global__ void k_rank_keys(PsView v, const int* __restrict__ new_elem, int ne, unsigned* keys, int* vals) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  const int e = m ? new_elem[s] : -1;
  keys[s] = e >= 0 ? (unsigned)e : (unsigned)ne;   // deleted / empty slots sort to the end
  vals[s] = s;
}
__global__ void k_rank_bounds(const unsigned* __restrict__ keys, int n, int ne, int* first, int* count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned e = keys[j];
  if (e >= (unsigned)ne) return;
  if (j == 0 || keys[j - 1] != e) first[e] = j;
  if (j == n - 1 || keys[j + 1] != e) count[e] = j + 1;   // end for now; turned into a count below
}
__global__ void k_rank_counts(const int* __restrict__ first, int* count, int ne) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne && count[e] > 0) count[e] -= first[e];
}
__global__ void k_rank_scatter(const unsigned* __restrict__ keys, const int* __restrict__ vals, int n, int ne,
                               const int* __restrict__ first, int* rank) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned e = keys[j];
  if (e < (unsigned)ne) rank[vals[j]] = j - first[e];
}
// row_fill[row(e)] = kept particles of e, so that new particles are appended behind them
__global__ void k_fill_from_kept(const int* __restrict__ kept, int ne, const int* __restrict__ elem2row, int* row_fill) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) row_fill[elem2row ? elem2row[e] : e] = kept[e];
}

// ------------------------------------------------------------------------------------------
// reshuffle (SCS_rebuild.h:4-120): when every row has at least as many holes as particles moving
// into it, only the movers travel (each into a hole of its destination row) and nothing else of
// the structure changes.  Same rules as the reference: a hole is a slot that holds no particle
// after the deletions (a mover's own slot is NOT a hole for this round).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int slot_row(const PsView& v, int slot) {
  int S = __ldg(v.tile_slice + (slot >> 5));
  while (slot >= __ldg(v.offsets + S + 1)) ++S;
  const int r = (slot - __ldg(v.offsets + S)) % v.C;
  return __ldg(v.slice_to_chunk + S) * v.C + r;
}
// incoming particles and holes per row; new_mask = particle after deletions
__global__ void k_shuffle_count(PsView v, const int* __restrict__ new_elem, const int* __restrict__ elem2row,
                                int* incoming, int* holes, int* outgoing, uint32_t* new_mask) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool is_particle = false;
  if (s < v.capacity) {
    const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
    const int row = slot_row(v, s);
    const int e = m ? new_elem[s] : -1;
    is_particle = m && e != -1;
    if (is_particle) {
      const int nrow = elem2row[e];
      if (nrow != row) { atomicAdd(incoming + nrow, 1); atomicAdd(outgoing + row, 1); }
    } else {
      atomicAdd(holes + row, 1);
      if (m) atomicAdd(outgoing + row, 1);      // deleted
    }
  }
  const unsigned w = __ballot_sync(0xffffffffu, is_particle);
  if ((threadIdx.x & 31) == 0 && s < v.capacity) new_mask[s >> 5] = w;
}
__global__ void k_shuffle_count_new(const int* __restrict__ elems, int n, const int* __restrict__ elem2row,
                                    int* incoming) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(incoming + elem2row[elems[i]], 1);
}
__global__ void k_shuffle_row_counts(int* row_ppe, const int* __restrict__ incoming,
                                     const int* __restrict__ outgoing, int nrows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrows) row_ppe[r] += incoming[r] - outgoing[r];
}
__global__ void k_shuffle_fits(const int* __restrict__ incoming, const int* __restrict__ holes, int nrows,
                               int* fail) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nrows && incoming[r] > holes[r]) *fail = 1;
}
// movers, grouped by destination row: src >= 0 is a slot of the structure, src < 0 is new particle -src-1
__global__ void k_shuffle_gather(PsView v, const int* __restrict__ new_elem, const int* __restrict__ elem2row,
                                 int* cursor, int* src) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (!m) return;
  const int e = new_elem[s];
  if (e == -1) return;
  const int nrow = elem2row[e];
  if (nrow != slot_row(v, s)) src[atomicAdd(cursor + nrow, 1)] = s;
}
__global__ void k_shuffle_gather_new(const int* __restrict__ elems, int n, const int* __restrict__ elem2row,
                                     int* cursor, int* src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) src[atomicAdd(cursor + elem2row[elems[i]], 1)] = -i - 1;
}
// every hole of row r takes the next mover bound for r, while there is one
__global__ void k_shuffle_holes(PsView v, const uint32_t* __restrict__ new_mask, int* next,
                                const int* __restrict__ end, int* hole) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  if ((new_mask[s >> 5] >> (s & 31)) & 1u) return;
  const int row = slot_row(v, s);
  if (next[row] >= end[row]) return;              // cheap pre-test, the claim below decides
  const int idx = atomicAdd(next + row, 1);
  if (idx < end[row]) hole[idx] = s;
}
__global__ void k_shuffle_move(const int* __restrict__ src, const int* __restrict__ hole, int nmove,
                               MemberTable mt, MemberTable mt_new, long stride, int n_new, uint32_t* mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmove) return;
  const int from = src[i], to = hole[i];
  const bool fresh = from < 0;
  const long fi = fresh ? -from - 1 : from;
  const MemberTable& t = fresh ? mt_new : mt;
  const long fstride = fresh ? n_new : stride;
  for (int k = 0; k < t.n; ++k) {
    const int sb = t.bytes[k];
    for (int c = 0; c < t.ncomp[k]; ++c) {
      const char* a = t.src[k] + ((long)c * fstride + fi) * sb;
      char* b = mt.dst[k] + ((long)c * stride + to) * sb;
      if (sb == 8) *(double*)b = *(const double*)a;
      else if (sb == 4) *(int*)b = *(const int*)a;
      else for (int q = 0; q < sb; ++q) b[q] = a[q];
    }
  }
  if (!fresh) atomicAnd(mask + (from >> 5), ~(1u << (from & 31)));
  atomicOr(mask + (to >> 5), 1u << (to & 31));
}

int g_rank_sort_ppe = 128;   // particles per element from which ranks come from a sort
int g_hist_block = 1;        // destination histogram with block-level reservation (0: one atomic per particle, A/B)
int g_staged_rebuild = 2;   // 2: single-pass gather (SCS, C = 32); 1: record stage; 0: direct scatter (A/B)
int g_sm_count_scs = 0;

pp_status stage_ensure(pp_ps* ps, size_t bytes, cudaStream_t s) {
  if (ps->stage_bytes >= bytes) return PP_OK;
  if (ps->stage) PP_CUDA(cudaFreeAsync(ps->stage, s));
  ps->stage = nullptr; ps->stage_bytes = 0;
  bytes += bytes / 8;                  // head room: capacity drifts by a few percent per rebuild
  PP_CUDA(cudaMallocAsync((void**)&ps->stage, bytes, s));
  ps->stage_bytes = bytes;
  return PP_OK;
}

pp_status scan_exclusive(const int* in, int* out, int n, cudaStream_t s) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s);
  char* tmp;
  PP_TRY(pp_dev_alloc(&tmp, tb, s));
  PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, s));
  pp_dev_free(tmp, s);
  return PP_OK;
}

void free_layout(ScsLayout& L, cudaStream_t s) {
  pp_dev_free(L.offsets, s); pp_dev_free(L.slice_to_chunk, s); pp_dev_free(L.row_to_element, s);
  pp_dev_free(L.element_to_row, s); pp_dev_free(L.chunk_start, s); pp_dev_free(L.tile_slice, s);
  pp_dev_free(L.row_ppe, s); pp_dev_free(L.mask, s);
  L = ScsLayout();
}

// Everything SellCSigma::construct derives from particles-per-element (SellCSigma.h:230-283)
pp_status scs_layout(const pp_ps_config& cfg, int ne, const int* ppe_dev, long np_bound, cudaStream_t s,
                     ScsLayout& L) {
  int* scal;
  double* inv;
  PP_TRY(pp_dev_alloc(&scal, 4, s));
  PP_TRY(pp_dev_alloc(&inv, 1, s));
  PP_CUDA(cudaMemsetAsync(scal, 0, 4 * sizeof(int), s));
  PP_CUDA(cudaMemsetAsync(inv, 0, sizeof(double), s));
  // chooseChunkHeight (SCS_buildFns.h:4-16)
  k_count_nonzero<<<std::min(pp_div_up(ne, kBlock), 1184), kBlock, 0, s>>>(ppe_dev, ne, scal + 2);
  int nnz = 0;
  PP_CUDA(cudaMemcpyAsync(&nnz, scal + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  const int Cmax = cfg.team_size > 0 ? cfg.team_size : 1;
  L.C = nnz == 0 ? 1 : (nnz < Cmax ? nnz : Cmax);
  const int C = L.C;
  L.nchunks = ne / C + (ne % C != 0);
  L.nrows = L.nchunks * C;
  // sigmaSort (SCS_sort.h:4-49, CUDA branch): ascending particle count inside windows of sigma
  // elements, stable.  Key = (window, count) packed into the fewest bits: counts are bounded by
  // the particle total, so a full sort of 1 M elements takes 3 radix passes instead of 8.
  int* sorted_elem = nullptr;
  if (cfg.sigma > 1) {
    const int sigma = cfg.sigma < ne ? cfg.sigma : ne;
    int cbits = 1;
    while (cbits < 31 && (1ll << cbits) <= (long long)np_bound) ++cbits;
    const int nwin = (ne + sigma - 1) / sigma;
    int wbits = 0;
    while ((1 << wbits) < nwin) ++wbits;
    uint64_t *k_in, *k_out;
    int *v_in;
    PP_TRY(pp_dev_alloc(&k_in, ne, s));
    PP_TRY(pp_dev_alloc(&k_out, ne, s));
    PP_TRY(pp_dev_alloc(&v_in, ne, s));
    PP_TRY(pp_dev_alloc(&sorted_elem, ne, s));
    k_sort_keys<uint64_t><<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(ppe_dev, ne, sigma > 0 ? sigma : 1, cbits, k_in, v_in);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s);
    char* tmp;
    PP_TRY(pp_dev_alloc(&tmp, tb, s));
    PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s));
    pp_dev_free(tmp, s); pp_dev_free(k_in, s); pp_dev_free(k_out, s); pp_dev_free(v_in, s);
  }
  PP_TRY(pp_dev_alloc(&L.row_to_element, L.nrows, s));
  PP_TRY(pp_dev_alloc(&L.element_to_row, L.nrows, s));
  PP_TRY(pp_dev_alloc(&L.row_ppe, L.nrows, s));
  k_rows<<<pp_div_up(L.nrows, kBlock), kBlock, 0, s>>>(sorted_elem, ppe_dev, ne, L.nrows,
                                                      L.row_to_element, L.element_to_row, L.row_ppe);
  pp_dev_free(sorted_elem, s);
  // constructChunks (SCS_buildFns.h:19-98)
  int *width, *spc, *slice_off;
  PP_TRY(pp_dev_alloc(&width, L.nchunks, s));
  PP_TRY(pp_dev_alloc(&spc, L.nchunks + 1, s));
  PP_TRY(pp_dev_alloc(&slice_off, L.nchunks + 1, s));
  k_chunk_widths<<<std::min(pp_div_up((long)L.nchunks * 32, kBlock), 1184), kBlock, 0, s>>>(L.row_ppe, L.nchunks, C, width, scal, inv);
  if (cfg.shuffle_padding > 0)
    k_pad_widths<<<pp_div_up(L.nchunks, kBlock), kBlock, 0, s>>>(width, L.nchunks, scal, inv,
                                                                 cfg.shuffle_padding, cfg.padding_strat);
  // constructOffsets (SCS_buildFns.h:115-153)
  const int V = cfg.V > 0 ? cfg.V : 1;
  k_slices_per_chunk<<<pp_div_up(L.nchunks + 1, kBlock), kBlock, 0, s>>>(width, L.nchunks, V, spc);
  PP_TRY(scan_exclusive(spc, slice_off, L.nchunks + 1, s));
  PP_CUDA(cudaMemcpyAsync(&L.nslices, slice_off + L.nchunks, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaMemcpyAsync(&L.nonempty_chunks, scal + 1, sizeof(int), cudaMemcpyDeviceToHost, s));   // k_chunk_widths
  PP_CUDA(cudaStreamSynchronize(s));
  int* slice_size;
  PP_TRY(pp_dev_alloc(&slice_size, L.nslices + 1, s));
  PP_TRY(pp_dev_alloc(&L.slice_to_chunk, L.nslices + 1, s));
  PP_TRY(pp_dev_alloc(&L.offsets, L.nslices + 2, s));
  PP_CUDA(cudaMemsetAsync(slice_size, 0, sizeof(int) * (L.nslices + 1), s));
  k_fill_slices<<<pp_div_up(L.nchunks, kBlock), kBlock, 0, s>>>(width, slice_off, L.nchunks, V, C,
                                                               L.slice_to_chunk, slice_size);
  PP_TRY(scan_exclusive(slice_size, L.offsets, L.nslices + 1, s));
  PP_CUDA(cudaMemcpyAsync(&L.capacity, L.offsets + L.nslices, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  PP_TRY(pp_dev_alloc(&L.chunk_start, L.nchunks + 1, s));
  k_chunk_start<<<pp_div_up(L.nchunks + 1, kBlock), kBlock, 0, s>>>(slice_off, L.offsets, L.nchunks,
                                                               L.capacity, L.chunk_start);
  const int ntiles = (L.capacity + 31) / 32;
  PP_TRY(pp_dev_alloc(&L.tile_slice, ntiles + 1, s));
  if (ntiles > 0)
    k_tile_slice<<<pp_div_up(ntiles, kBlock), kBlock, 0, s>>>(L.offsets, L.nslices, ntiles, L.tile_slice);
  L.mask_words = ntiles + 1;
  PP_TRY(pp_dev_alloc(&L.mask, L.mask_words, s));
  PP_CUDA(cudaMemsetAsync(L.mask, 0, sizeof(uint32_t) * L.mask_words, s));
  PP_KERNEL_CHECK();
  pp_dev_free(width, s); pp_dev_free(spc, s); pp_dev_free(slice_off, s); pp_dev_free(slice_size, s);
  pp_dev_free(scal, s); pp_dev_free(inv, s);
  return PP_OK;
}

PsView layout_view(const ScsLayout& L, int ne) {
  PsView v;
  v.kind = PP_PS_SCS; v.capacity = L.capacity; v.mask_bits = L.mask; v.slot_elem = nullptr;
  v.offsets = L.offsets; v.slice_to_chunk = L.slice_to_chunk; v.row_to_element = L.row_to_element;
  v.tile_slice = L.tile_slice; v.C = L.C; v.nslices = L.nslices;
  v.chunk_start = L.chunk_start; v.nchunks = L.nchunks; v.nelems = ne;
  return v;
}

void adopt_layout(pp_ps* ps, ScsLayout& L, cudaStream_t s) {
  pp_dev_free(ps->offsets, s); pp_dev_free(ps->slice_to_chunk, s); pp_dev_free(ps->row_to_element, s);
  pp_dev_free(ps->element_to_row, s); pp_dev_free(ps->tile_slice, s); pp_dev_free(ps->mask_bits, s);
  pp_dev_free(ps->chunk_start, s); pp_dev_free(ps->row_ppe, s);
  ps->C = L.C; ps->nchunks = L.nchunks; ps->nrows = L.nrows; ps->nslices = L.nslices;
  ps->capacity = L.capacity;
  ps->offsets = L.offsets; ps->slice_to_chunk = L.slice_to_chunk;
  ps->row_to_element = L.row_to_element; ps->element_to_row = L.element_to_row;
  ps->tile_slice = L.tile_slice; ps->mask_bits = L.mask; ps->mask_words_alloc = L.mask_words;
  ps->chunk_start = L.chunk_start; ps->row_ppe = L.row_ppe;
  // Empty chunks have no slice: "more slices than chunks" would miss a wide chunk in a structure that also
  // has empty ones (pseudoXGCm's load: 47 K slices, 63 K chunks, one chunk of 924 slices).
  ps->sliced = L.nslices > (L.nonempty_chunks >= 0 ? L.nonempty_chunks : L.nchunks) ? 1 : 0;
  L = ScsLayout();
  if (ps->slot_elem) { pp_dev_free(ps->slot_elem, s); ps->slot_elem = nullptr; }
  ps->slot_elem_valid = false;
  ps->slot_elem_materialized = false;
  ps->first_chunk = 0;
}

pp_status member_table(const pp_ps* ps, const std::vector<void*>& src, const std::vector<void*>& dst,
                       MemberTable& mt) {
  PP_REQUIRE(ps->nmembers <= 16, "at most 16 particle members are supported");
  mt.n = ps->nmembers;
  for (int i = 0; i < ps->nmembers; ++i) {
    mt.src[i] = (char*)(src.empty() ? nullptr : src[i]);
    mt.dst[i] = (char*)dst[i];
    mt.bytes[i] = ps->members[i].scalar_bytes;
    mt.ncomp[i] = ps->members[i].ncomp;
  }
  return PP_OK;
}
}  // namespace

int g_try_shuffling = 1;

// SCS_rebuild.h:4-120.  done = true iff the particles fitted and the structure was updated in place.
pp_status try_reshuffle(pp_ps* ps, const int* new_element, int n_new, const int* new_particle_elements,
                        const void* const* new_particle_info, const MemberTable& mt_new_in, cudaStream_t s,
                        bool& done) {
  done = false;
  const int cap = ps->capacity, nrows = ps->nrows;
  const PsView v = ps->view();
  int *incoming, *holes, *outgoing, *fail;
  uint32_t* new_mask;
  const long nwords = (cap + 31) / 32;
  PP_TRY(pp_dev_alloc(&incoming, nrows + 1, s));
  PP_TRY(pp_dev_alloc(&holes, nrows + 1, s));
  PP_TRY(pp_dev_alloc(&outgoing, nrows + 1, s));
  PP_TRY(pp_dev_alloc(&fail, 2, s));
  PP_TRY(pp_dev_alloc(&new_mask, nwords + 1, s));
  PP_CUDA(cudaMemsetAsync(incoming, 0, sizeof(int) * (nrows + 1), s));
  PP_CUDA(cudaMemsetAsync(holes, 0, sizeof(int) * (nrows + 1), s));
  PP_CUDA(cudaMemsetAsync(outgoing, 0, sizeof(int) * (nrows + 1), s));
  PP_CUDA(cudaMemsetAsync(fail, 0, 2 * sizeof(int), s));
  k_shuffle_count<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(v, new_element, ps->element_to_row, incoming, holes,
                                                           outgoing, new_mask);
  if (n_new > 0)
    k_shuffle_count_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new,
                                                                    ps->element_to_row, incoming);
  k_shuffle_fits<<<pp_div_up(nrows, kBlock), kBlock, 0, s>>>(incoming, holes, nrows, fail);
  int* begin;
  PP_TRY(pp_dev_alloc(&begin, nrows + 2, s));
  PP_TRY(scan_exclusive(incoming, begin, nrows + 1, s));    // begin[nrows] = number of movers
  int h_fail = 0, nmove = 0;
  PP_CUDA(cudaMemcpyAsync(&h_fail, fail, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaMemcpyAsync(&nmove, begin + nrows, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  if (!h_fail) {
    // the deletions take effect; movers keep their bit until they have been copied
    PP_CUDA(cudaMemcpyAsync(ps->mask_bits, new_mask, sizeof(uint32_t) * nwords, cudaMemcpyDeviceToDevice, s));
    if (nmove > 0) {
      int *cursor, *src, *hole;
      PP_TRY(pp_dev_alloc(&cursor, nrows + 1, s));
      PP_TRY(pp_dev_alloc(&src, nmove, s));
      PP_TRY(pp_dev_alloc(&hole, nmove, s));
      PP_CUDA(cudaMemcpyAsync(cursor, begin, sizeof(int) * (nrows + 1), cudaMemcpyDeviceToDevice, s));
      k_shuffle_gather<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(v, new_element, ps->element_to_row, cursor, src);
      if (n_new > 0)
        k_shuffle_gather_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new,
                                                                         ps->element_to_row, cursor, src);
      // cursor[r] is now the end of row r's movers; begin[r] counts up as holes claim them
      k_shuffle_holes<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(v, new_mask, begin, cursor, hole);
      MemberTable mt, mtn = mt_new_in;
      PP_TRY(member_table(ps, ps->data, ps->data, mt));
      (void)new_particle_info;
      k_shuffle_move<<<pp_div_up(nmove, kBlock), kBlock, 0, s>>>(src, hole, nmove, mt, mtn, ps->stride, n_new,
                                                                 ps->mask_bits);
      pp_dev_free(cursor, s); pp_dev_free(src, s); pp_dev_free(hole, s);
    }
    k_shuffle_row_counts<<<pp_div_up(nrows, kBlock), kBlock, 0, s>>>(ps->row_ppe, incoming, outgoing, nrows);
    PP_KERNEL_CHECK();
    done = true;
  }
  pp_dev_free(outgoing, s);
  pp_dev_free(incoming, s); pp_dev_free(holes, s); pp_dev_free(fail, s); pp_dev_free(new_mask, s);
  pp_dev_free(begin, s);
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// SellCSigma::construct (SellCSigma.h:230-283)
// ------------------------------------------------------------------------------------------
pp_status pp_scs_build(pp_ps* ps, const int* ppe_dev, const int* pelems_dev,
                       const void* const* pinfo, int memspace, cudaStream_t s) {
  ScsLayout L;
  PP_TRY(scs_layout(ps->cfg, ps->nelems, ppe_dev, 0x7fffffffL, s, L));   // caller-provided counts: no tighter bound
  const int ne = ps->nelems, np = ps->nptcls;
  if (L.capacity > 0 && np > 0) {
    PsView v = layout_view(L, ne);
    k_scs_mask<<<pp_div_up(L.capacity, kBlock), kBlock, 0, s>>>(v, L.row_ppe, ne, L.mask, L.mask_words);
  }
  adopt_layout(ps, L, s);
  ps->V = ps->cfg.V;
  // allocate the data and its swap copy with extra padding (SellCSigma.h:264-272)
  long cap = ps->capacity;
  if (ps->cfg.extra_padding > 0) cap = (long)(cap * (1 + ps->cfg.extra_padding));
  if (cap < 1) cap = 1;
  ps->stride = cap;
  ps->data_alloc = cap;
  PP_TRY(pp_ps_alloc_members(ps, ps->data, ps->stride, s));
  if (!ps->cfg.always_realloc) {
    PP_TRY(pp_ps_alloc_members(ps, ps->swap, ps->stride, s));
    ps->swap_stride = ps->stride;
    ps->swap_alloc = ps->stride;
  }
  // initSCSData (SCS_buildFns.h:202-225)
  if (np > 0 && pelems_dev && pinfo) {
    int *row_fill, *slots;
    PP_TRY(pp_dev_alloc(&row_fill, ps->nrows, s));
    PP_TRY(pp_dev_alloc(&slots, np, s));
    PP_CUDA(cudaMemsetAsync(row_fill, 0, sizeof(int) * ps->nrows, s));
    k_assign_slots<<<pp_div_up(np, kBlock), kBlock, 0, s>>>(pelems_dev, np, ps->element_to_row,
                                                            ps->chunk_start, ps->C, row_fill, slots);
    MemberTable mt;
    std::vector<void*> srcs(ps->nmembers, nullptr);
    std::vector<char*> staged(ps->nmembers, nullptr);
    for (int i = 0; i < ps->nmembers; ++i) {
      const size_t bytes = (size_t)ps->members[i].scalar_bytes * ps->members[i].ncomp * np;
      PP_TRY(pp_dev_import(&staged[i], (const char*)pinfo[i], bytes, memspace, s));
      srcs[i] = staged[i];
    }
    PP_TRY(member_table(ps, srcs, ps->data, mt));
    k_place_new<<<pp_div_up(np, kBlock), kBlock, 0, s>>>(slots, np, mt, ps->stride);
    for (char* p : staged) pp_dev_free(p, s);
    pp_dev_free(row_fill, s); pp_dev_free(slots, s);
  }
  PP_KERNEL_CHECK();
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// Single-pass rebuild of a Sell-C-sigma structure (C = 32): device-side layout + gather.
//
// The staged move above costs 2 x record (SoA) + 2 x padded record (stage) of DRAM traffic and
// the layout code around it reads five scalars back to the host.  This path
//   * derives the whole layout on the device with ONE host read at the end (active particles,
//     non-empty elements, largest row, slices, capacity), speculating on the two values the host
//     needs earlier: the chunk height (C = team size unless fewer than C elements hold particles,
//     SCS_buildFns.h:4-16) and the number of key bits of the row sort (from the previous rebuild's
//     largest row).  A failed speculation falls back to the synchronous path above;
//   * moves every record ONCE: the destination slot of each kept particle follows from its rank
//     in its element (the value the histogram's atomic returned), an inverse map
//     src_of[destination] = source is written (4-byte scatter into an L2-resident array), and one
//     warp per destination chunk gathers the records component by component: destination stores
//     are fully coalesced, source loads are 8-byte gathers;
//   * hands the chunks to the warps in ascending order of their first row's element id, not in
//     slot order.  Rows are sorted by particle count, so slot order sweeps the element range
//     once per count class and the other three particles of a gathered 32-byte source sector
//     would be needed a whole sweep later; in element order they are needed by chunks that are
//     resident at the same time, and the sector is served from L2.
// ------------------------------------------------------------------------------------------
// key = 1/256-th of the element range the chunk's first row lies in: one radix pass; chunks of one
// bucket (~4 K elements) are in flight together anyway
// empty chunks (width 0; on a PICpart that buffers the whole mesh most of them) sort behind all others:
// the gather only takes the non-empty ones.  nb buckets of the element range (+ 1 for the empty chunks).
__global__ void k_chunk_keys(const int* __restrict__ row2elem, const int* __restrict__ width, int nchunks, int clo,
                             unsigned nb, unsigned* keys, int* vals) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = clo + k;
  if (c >= nchunks) return;
  const unsigned b = (unsigned)(((long long)row2elem[c * 32] * nb) / ((long long)nchunks * 32));
  keys[k] = (width && width[c] == 0) ? nb : (b < nb - 1 ? b : nb - 1);
  vals[k] = c;
}
// chunks of a C = 32 layout in ascending order of (the bucket of) their first row's element
// clo: chunks below it are known to be empty (a mostly empty structure: the caller's bound on the rows
// that hold particles); only chunks [clo, nchunks) are ordered, the non-empty ones first
pp_status chunk_order_build(const int* row_to_element, const int* width, int nchunks, cudaStream_t s, int** order,
                            int clo = 0) {
  *order = nullptr;
  if (nchunks < 2) return PP_OK;
  const int n = nchunks - clo;
  unsigned *ck_in, *ck_out;
  int* cv_in;
  PP_TRY(pp_dev_alloc(&ck_in, n, s)); PP_TRY(pp_dev_alloc(&ck_out, n, s));
  PP_TRY(pp_dev_alloc(&cv_in, n, s)); PP_TRY(pp_dev_alloc(order, n, s));
  // one radix pass (255 buckets) for structures of up to 64 K chunks, two passes (65535 buckets) above:
  // on a full-mesh PICpart the particles sit in a small part of the element range
  const int ebits = nchunks > 65536 ? 16 : 8;
  k_chunk_keys<<<pp_div_up(n, kBlock), kBlock, 0, s>>>(row_to_element, width, nchunks, clo, (1u << ebits) - 1u,
                                                       ck_in, cv_in);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, ck_in, ck_out, cv_in, *order, n, 0, ebits, s);
  char* tmp;
  PP_TRY(pp_dev_alloc(&tmp, tb, s));
  PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, ck_in, ck_out, cv_in, *order, n, 0, ebits, s));
  pp_dev_free(tmp, s); pp_dev_free(ck_in, s); pp_dev_free(ck_out, s); pp_dev_free(cv_in, s);
  return PP_OK;
}

// (Re)size the swap copy of the member arrays for a structure of `capacity` slots.  The condition
// is SCS_rebuild.h:223-229 as written; with the default minimize_size = 0.8 it holds on nearly every
// rebuild, and the reference then frees and re-creates (zero-filled) views of capacity * (1 +
// extra_padding) slots.  The sizes follow the reference; the memory is only exchanged when the
// existing allocation is too small: otherwise the same arrays are reinterpreted with the new stride
// and not cleared (slots that hold no particle keep stale bytes instead of zeros; nothing reads them).
pp_status ensure_swap(pp_ps* ps, long capacity, cudaStream_t s) {
  const pp_ps_config& cfg = ps->cfg;
  if (!(cfg.always_realloc || ps->swap.empty() || ps->swap_stride < capacity ||
        ps->swap_stride * cfg.minimize_size < capacity))
    return PP_OK;
  long nstride = (long)(capacity * (1 + cfg.extra_padding));
  if (nstride < capacity) nstride = capacity;
  if (nstride < 1) nstride = 1;
  if (!cfg.always_realloc && !ps->swap.empty() && ps->swap_alloc >= nstride) {
    ps->swap_stride = nstride;
    return PP_OK;
  }
  for (void* p : ps->swap) pp_dev_free((char*)p, s);
  PP_TRY(pp_ps_alloc_members(ps, ps->swap, nstride, s));
  ps->swap_stride = nstride;
  ps->swap_alloc = nstride;
  return PP_OK;
}

struct FastScal {        // device scalars of one rebuild, read back once
  int nnz, active, maxcount, bad;
  int cw_sum, cw_cnt, nslices, capacity;
  int next_chunk, pad0, pad1, pad2;
  double inv;
};

__global__ void k_count_stats(const int* __restrict__ a, int n, FastScal* out) {
  __shared__ int p_nz[kBlock / 32], p_sum[kBlock / 32], p_max[kBlock / 32];
  int nz = 0, sum = 0, mx = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int v = a[i];
    nz += v > 0; sum += v; mx = max(mx, v);
  }
  nz = __reduce_add_sync(0xffffffffu, nz);
  sum = __reduce_add_sync(0xffffffffu, sum);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) { p_nz[threadIdx.x >> 5] = nz; p_sum[threadIdx.x >> 5] = sum; p_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { nz += p_nz[w]; sum += p_sum[w]; mx = max(mx, p_max[w]); }
    if (nz) atomicAdd(&out->nnz, nz);
    if (sum) atomicAdd(&out->active, sum);
    if (mx) atomicMax(&out->maxcount, mx);
  }
}

// k_count_stats + k_sort_keys in one pass over the counts
template <class Key>
__global__ void k_stats_keys(const int* __restrict__ a, int n, int sigma, int cbits, Key* keys, int* vals,
                             FastScal* out) {
  __shared__ int p_nz[kBlock / 32], p_sum[kBlock / 32], p_max[kBlock / 32];
  int nz = 0, sum = 0, mx = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int v = a[i];
    nz += v > 0; sum += v; mx = max(mx, v);
    const Key win = (Key)(i / sigma);
    keys[i] = (win << cbits) | (Key)(uint32_t)v;
    vals[i] = (int)i;
  }
  nz = __reduce_add_sync(0xffffffffu, nz);
  sum = __reduce_add_sync(0xffffffffu, sum);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) { p_nz[threadIdx.x >> 5] = nz; p_sum[threadIdx.x >> 5] = sum; p_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { nz += p_nz[w]; sum += p_sum[w]; mx = max(mx, p_max[w]); }
    if (nz) atomicAdd(&out->nnz, nz);
    if (sum) atomicAdd(&out->active, sum);
    if (mx) atomicMax(&out->maxcount, mx);
  }
}
// k_rows + k_chunk_widths for C = 32: one warp per chunk, lane = row
__global__ void k_rows_widths(const int* __restrict__ sorted_elem, const int* __restrict__ ppe, int ne, int nchunks,
                              int* row2elem, int* elem2row, int* row_ppe, int* width, int* cw, double* inv) {
  __shared__ int p_sum[kBlock / 32], p_cnt[kBlock / 32];
  __shared__ double p_inv[kBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int sum = 0, cnt = 0;
  double isum = 0.0;
  for (long c = blockIdx.x * (long)(kBlock / 32) + wid; c < nchunks; c += (long)gridDim.x * (kBlock / 32)) {
    const int i = (int)c * 32 + lane;
    int np = 0;
    if (i < ne) {
      const int e = sorted_elem ? sorted_elem[i] : i;
      np = ppe[e];
      row2elem[i] = e; elem2row[e] = i; row_ppe[i] = np;
    } else {               // padding rows up to a multiple of C (SCS_buildFns.h:39-44)
      row2elem[i] = i; elem2row[i] = i; row_ppe[i] = 0;
    }
    const int w = __reduce_max_sync(0xffffffffu, np);
    if (lane == 0) {
      width[c] = w;
      if (w > 0) { sum += w; cnt += 1; isum += 1.0 / w; }
    }
  }
  if (lane == 0) { p_sum[wid] = sum; p_cnt[wid] = cnt; p_inv[wid] = isum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { sum += p_sum[w]; cnt += p_cnt[w]; isum += p_inv[w]; }
    if (cnt) { atomicAdd(cw, sum); atomicAdd(cw + 1, cnt); atomicAdd(inv, isum); }
  }
}

// ---- mostly empty structures (a PICpart that buffers the whole mesh holds particles in its own share of
// the rows only): the stable ascending sort of ONE window puts the empty rows first, in element order,
// and the others behind them by count.  The empty rows need no sort: row = number of empty elements in
// front (from a prefix sum of the non-empty flags); only the non-empty ones, compacted, are sorted.
struct NonZeroFlag {
  __host__ __device__ int operator()(const int& v) const { return v > 0 ? 1 : 0; }
};
// pos[i] = non-empty elements in front of i.  Empty element: its row arrays are written here.
// Non-empty element: (count, element) goes to entry pos[i] of the sort's input (if the bound m holds).
__global__ void k_split_rows(const int* __restrict__ a, const int* __restrict__ pos, int n, int m, uint32_t* keys,
                             int* vals, int* row2elem, int* elem2row, int* row_ppe, FastScal* out) {
  __shared__ int p_nz[kBlock / 32], p_sum[kBlock / 32], p_max[kBlock / 32];
  int nz = 0, sum = 0, mx = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int v = a[i];
    const int p = pos[i];
    nz += v > 0; sum += v; mx = max(mx, v);
    if (v > 0) {
      if (p < m) { keys[p] = (uint32_t)v; vals[p] = (int)i; }
    } else {
      const int row = (int)i - p;
      row2elem[row] = (int)i; elem2row[i] = row; row_ppe[row] = 0;
    }
  }
  nz = __reduce_add_sync(0xffffffffu, nz);
  sum = __reduce_add_sync(0xffffffffu, sum);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) { p_nz[threadIdx.x >> 5] = nz; p_sum[threadIdx.x >> 5] = sum; p_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { nz += p_nz[w]; sum += p_sum[w]; mx = max(mx, p_max[w]); }
    if (nz) atomicAdd(&out->nnz, nz);
    if (sum) atomicAdd(&out->active, sum);
    if (mx) atomicMax(&out->maxcount, mx);
  }
}
// k_rows_widths for the split layout: rows [0, z) are the empty ones (written by k_split_rows), row z + k is
// entry (m - nnz) + k of the sorted non-empty elements (the first m - nnz entries are key-0 fillers)
__global__ void k_rows_widths_split(const int* __restrict__ sorted_nz, const int* __restrict__ ppe, int ne, int nchunks,
                                    int m, const FastScal* __restrict__ sc, int* row2elem, int* elem2row, int* row_ppe,
                                    int* width, int* cw, double* inv) {
  __shared__ int p_sum[kBlock / 32], p_cnt[kBlock / 32];
  __shared__ double p_inv[kBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nnz = sc->nnz;
  const bool good = nnz <= m;                  // else the host drops this layout (speculation failed)
  const int z = ne - nnz;
  int sum = 0, cnt = 0;
  double isum = 0.0;
  for (long c = blockIdx.x * (long)(kBlock / 32) + wid; c < nchunks; c += (long)gridDim.x * (kBlock / 32)) {
    const int i = (int)c * 32 + lane;
    int np = 0;
    if (!good || (int)c * 32 + 31 < z) {       // a chunk of empty rows: nothing to read
      if (lane == 0) width[c] = 0;
      continue;
    }
    if (i >= ne) {                             // padding rows up to a multiple of C (SCS_buildFns.h:39-44)
      row2elem[i] = i; elem2row[i] = i; row_ppe[i] = 0;
    } else if (i >= z) {
      const int e = sorted_nz[(i - z) + (m - nnz)];
      np = ppe[e];
      row2elem[i] = e; elem2row[e] = i; row_ppe[i] = np;
    }
    const int w = __reduce_max_sync(0xffffffffu, np);
    if (lane == 0) {
      width[c] = w;
      if (w > 0) { sum += w; cnt += 1; isum += 1.0 / w; }
    }
  }
  if (lane == 0) { p_sum[wid] = sum; p_cnt[wid] = cnt; p_inv[wid] = isum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { sum += p_sum[w]; cnt += p_cnt[w]; isum += p_inv[w]; }
    if (cnt) { atomicAdd(cw, sum); atomicAdd(cw + 1, cnt); atomicAdd(inv, isum); }
  }
}

// padding (SCS_buildFns.h:62-97) + slices per chunk + slots per chunk in one pass
__global__ void k_chunk_sizes(int* width, int nchunks, const FastScal* __restrict__ sc, double pad, int strat,
                              int V, int C, int2* sizes) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nchunks) return;
  if (c == nchunks) { sizes[c] = make_int2(0, 0); return; }
  int w = width[c];
  if (pad > 0 && sc->cw_sum > 0) {
    if (strat == PP_PAD_EVENLY) {
      const int avg_pad = (int)(sc->cw_sum * pad / sc->cw_cnt);
      if (w > 0) w = w + avg_pad;
    } else if (strat == PP_PAD_PROPORTIONALLY) {
      w = (int)(w + w * pad);
    } else {
      const double cw_sum2 = sc->cw_sum / sc->inv * pad;
      if (w != 0) w = (int)(w + cw_sum2 / w);
    }
    width[c] = w;
  }
  sizes[c] = make_int2(w / V + (w % V != 0), w * C);
}
struct Int2Sum {
  __host__ __device__ int2 operator()(const int2& a, const int2& b) const { return make_int2(a.x + b.x, a.y + b.y); }
};
// constructOffsets (SCS_buildFns.h:115-153) from the two prefix sums
__global__ void k_fill_layout(const int2* __restrict__ pref, const int* __restrict__ width, int nchunks, int V,
                              int C, int* s2c, int* offsets, int* chunk_start, FastScal* sc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nchunks) return;
  const int2 p = pref[c];
  chunk_start[c] = p.y;
  if (c == nchunks) { offsets[p.x] = p.y; offsets[p.x + 1] = p.y; sc->nslices = p.x; sc->capacity = p.y; return; }
  const int ns = pref[c + 1].x - p.x;
  for (int j = 0; j < ns; ++j) { s2c[p.x + j] = c; offsets[p.x + j] = p.y + j * V * C; }
}
// kept particles: src_of[destination slot] = source slot
__global__ void k_invmap(PsView v, const int* __restrict__ new_elem, const int* __restrict__ rank,
                         const int* __restrict__ elem2row, const int* __restrict__ chunk_start, int* src_of) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  const bool m = (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
  if (!m) return;
  const int e = __ldg(new_elem + s);
  if (e < 0) return;
  const int row = __ldg(elem2row + e);
  src_of[__ldg(chunk_start + (row >> 5)) + (__ldg(rank + s) << 5) + (row & 31)] = s;
}

// One block per destination chunk, the chunk's columns dealt round-robin to the block's warps
// (lane = row): gathers the records of the chunk's particles into the new SoA columns and writes
// the chunk's particle mask.  Few chunks are in flight (one per resident block), each with many
// loads outstanding (a warp works on two columns at a time): the source sectors touched by the
// chunks in flight must stay L2-resident until the chunks that own their other three particles
// have read them too.
// t: old structure -> new structure; tn: arrays of the particles being added -> new structure.
__device__ __forceinline__ void gather_new(const UnitTable& tn, long slot, int src) {
  const long i = -(long)src - 1;
  for (int u = 0; u < tn.nunits; ++u) unit_store(tn, u, slot, unit_load(tn, u, i));
}
template <int kGatherWarps>
__global__ void __launch_bounds__(kGatherWarps * 32) k_gather_scs(
    const int* __restrict__ chunk_start, const int* __restrict__ row_ppe, const int* __restrict__ order,
    int nchunks, const __grid_constant__ UnitTable t, const __grid_constant__ UnitTable tn,
    const int* __restrict__ src_of, uint32_t* mask, int* next_chunk) {
  __shared__ int s_k[2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  constexpr int U = 10;
  if (threadIdx.x == 0) s_k[0] = atomicAdd(next_chunk, 1);
  for (int round = 0;; ++round) {
    __syncthreads();                                   // s_k[round & 1] is published; the other entry is free
    const int k = s_k[round & 1];
    if (k >= nchunks) break;
    if (threadIdx.x == 0) s_k[(round + 1) & 1] = atomicAdd(next_chunk, 1);   // overlaps this chunk's work
    const int c = order ? __ldg(order + k) : k;
    const int cs = __ldg(chunk_start + c);
    const int ncols = (__ldg(chunk_start + c + 1) - cs) >> 5;
    const int ppe = __ldg(row_ppe + c * 32 + lane);
    for (int colA = wid; colA < ncols; colA += 2 * kGatherWarps) {
      const int colB = colA + kGatherWarps;
      const bool vA = colA < ppe, vB = colB < ncols && colB < ppe;
      const unsigned wA = __ballot_sync(full, vA), wB = __ballot_sync(full, vB);
      if (lane == 0) {
        mask[(cs >> 5) + colA] = wA;
        if (colB < ncols) mask[(cs >> 5) + colB] = wB;
      }
      const long slotA = (long)cs + colA * 32 + lane, slotB = (long)cs + colB * 32 + lane;
      const int srcA = vA ? __ldg(src_of + slotA) : -1, srcB = vB ? __ldg(src_of + slotB) : -1;
      const bool oA = vA && srcA >= 0, oB = vB && srcB >= 0;      // from the old structure
      for (int u0 = 0; u0 < t.nunits; u0 += U) {
        unsigned long long a[U], b[U];
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (u0 + q < t.nunits && oA) a[q] = unit_load(t, u0 + q, srcA);
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (u0 + q < t.nunits && oB) b[q] = unit_load(t, u0 + q, srcB);
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (u0 + q < t.nunits && oA) unit_store(t, u0 + q, slotA, a[q]);
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (u0 + q < t.nunits && oB) unit_store(t, u0 + q, slotB, b[q]);
      }
      if (vA && !oA) gather_new(tn, slotA, srcA);
      if (vB && !oB) gather_new(tn, slotB, srcB);
    }
  }
}

int g_rebuild_chunk_order = 1;   // 0: chunks in slot order (A/B)
int g_rebuild_split_rows = 1;    // 0: always sort all rows (A/B)
int g_gather_bps = 0;            // resident blocks per SM of the gather; 0 = from the chunks' footprint
double g_gather_l2_bytes = 48e6; // footprint the chunks in flight may have
double g_gather_max_cols = 14.0; // average columns per chunk up to which the gather beats the record stage

// returns done = false when a speculation failed (nothing of the structure has changed then)
pp_status rebuild_scs_gather(pp_ps* ps, const int* new_element, int n_new, const int* n_new_dev, long new_ld,
                             const int* new_particle_elements, const void* const* new_particle_info,
                             const int* remap, cudaStream_t s, bool& done) {
  done = false;
  const int ne = ps->nelems, C = 32, cap = ps->capacity;
  const pp_ps_config& cfg = ps->cfg;
  const int nchunks = ne / C + (ne % C != 0), nrows = nchunks * C;
  const int V = cfg.V > 0 ? cfg.V : 1;
  const long active_bound = (long)ps->nptcls + n_new;
  // key bits of the row sort: speculate from the largest row of the previous rebuild
  int cbits = 1;
  while (cbits < 31 && (1ll << cbits) <= (long long)active_bound) ++cbits;
  if (ps->ppe_bits_hint > 0 && ps->ppe_bits_hint < cbits) cbits = ps->ppe_bits_hint;
  FastScal* sc;
  PP_TRY(pp_dev_alloc(&sc, 1, s));
  PP_CUDA(cudaMemsetAsync(sc, 0, sizeof(FastScal), s));
  int *count, *rank = nullptr, *kept = nullptr;
  PP_TRY(pp_dev_alloc(&count, ne + 1, s));
  PP_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ne + 1), s));
  int* rank_new = nullptr;
  // new-particle kernels: grid-stride; a device-side count (n_new is only a bound then) gets a fixed grid
  const int new_grid = n_new_dev ? std::min(pp_div_up(n_new, kBlock), 1184) : pp_div_up(n_new, kBlock);
  {
    PP_TIME_KIND(s, ps->cfg.kind, "count active particles");       // SCS_rebuild.h:133-166
    if (cap > 0) {
      PP_TRY(pp_dev_alloc(&rank, cap, s));
      if (g_hist_block) k_hist_kept_block<<<pp_div_up(cap, kHistSlots), 256, 0, s>>>(ps->view(), new_element, count, rank);
      else k_hist_kept<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(ps->view(), new_element, count, rank);
    }
    if (n_new > 0) {               // counted after the kept particles: their ranks follow the kept ones
      PP_TRY(pp_dev_alloc(&rank_new, n_new, s));
      k_hist_new<<<new_grid, kBlock, 0, s>>>(new_particle_elements, n_new, count, &sc->bad, n_new_dev, rank_new);
    }
    if (cfg.sigma <= 1) k_count_stats<<<std::min(pp_div_up(ne, kBlock), 592), kBlock, 0, s>>>(count, ne, sc);
  }
  PPTimeScope* t_build = new PPTimeScope(s, (std::string(pp_kind_name(ps->cfg.kind)) + " SCS specific building").c_str());
  // ---- layout (scs_layout above, without its host reads)
  ScsLayout L;
  L.C = C; L.nchunks = nchunks; L.nrows = nrows;
  int* sorted_elem = nullptr;
  PP_TRY(pp_dev_alloc(&L.row_to_element, nrows, s));
  PP_TRY(pp_dev_alloc(&L.element_to_row, nrows, s));
  PP_TRY(pp_dev_alloc(&L.row_ppe, nrows, s));
  int* width;
  int2 *sizes, *pref;
  PP_TRY(pp_dev_alloc(&width, nchunks, s));
  PP_TRY(pp_dev_alloc(&sizes, nchunks + 1, s));
  PP_TRY(pp_dev_alloc(&pref, nchunks + 1, s));
  // A mostly empty structure (one sort window; the previous rebuild left fewer than 40 % of the rows
  // non-empty): only the non-empty rows are sorted, see k_split_rows.  m bounds their number -- a
  // speculation like the key width, checked in the one host read.
  const long m_want = (long)ps->nnz_hint + ps->nnz_hint / 4 + 4096;
  const bool split = g_rebuild_split_rows && cfg.sigma >= ne && ps->nnz_hint > 0 && m_want * 2 < ne && cbits <= 31;
  const int m_split = split ? (int)m_want : 0;
  if (split) {
    int *pos, *v_in, *sorted_nz;
    uint32_t *k_in, *k_out;
    PP_TRY(pp_dev_alloc(&pos, ne, s));
    PP_TRY(pp_dev_alloc(&k_in, m_split, s)); PP_TRY(pp_dev_alloc(&k_out, m_split, s));
    PP_TRY(pp_dev_alloc(&v_in, m_split, s)); PP_TRY(pp_dev_alloc(&sorted_nz, m_split, s));
    PP_CUDA(cudaMemsetAsync(k_in, 0, sizeof(uint32_t) * m_split, s));     // fillers: key 0 sorts first
    cub::TransformInputIterator<int, NonZeroFlag, const int*> flags(count, NonZeroFlag());
    size_t tb = 0, tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, pos, ne, s);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, k_in, k_out, v_in, sorted_nz, m_split, 0, cbits, s);
    char* tmp;
    PP_TRY(pp_dev_alloc(&tmp, tb > tb2 ? tb : tb2, s));
    PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, flags, pos, ne, s));
    k_split_rows<<<std::min(pp_div_up(ne, kBlock), 2368), kBlock, 0, s>>>(count, pos, ne, m_split, k_in, v_in,
                                                                          L.row_to_element, L.element_to_row, L.row_ppe, sc);
    PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb2, k_in, k_out, v_in, sorted_nz, m_split, 0, cbits, s));
    k_rows_widths_split<<<std::min(pp_div_up((long)nchunks * 32, kBlock), 2368), kBlock, 0, s>>>(
        sorted_nz, count, ne, nchunks, m_split, sc, L.row_to_element, L.element_to_row, L.row_ppe, width, &sc->cw_sum,
        &sc->inv);
    pp_dev_free(tmp, s); pp_dev_free(pos, s); pp_dev_free(k_in, s); pp_dev_free(k_out, s);
    pp_dev_free(v_in, s); pp_dev_free(sorted_nz, s);
  } else if (cfg.sigma > 1) {
    const int sigma = cfg.sigma < ne ? cfg.sigma : ne;
    const int nwin = (ne + sigma - 1) / sigma;
    int wbits = 0;
    while ((1 << wbits) < nwin) ++wbits;
    int* v_in;
    PP_TRY(pp_dev_alloc(&v_in, ne, s)); PP_TRY(pp_dev_alloc(&sorted_elem, ne, s));
    char* tmp;
    size_t tb = 0;
    if (cbits + wbits <= 32) {       // the usual case (one window, or few): half the key traffic
      uint32_t *k_in, *k_out;
      PP_TRY(pp_dev_alloc(&k_in, ne, s)); PP_TRY(pp_dev_alloc(&k_out, ne, s));
      k_stats_keys<uint32_t><<<std::min(pp_div_up(ne, kBlock), 1184), kBlock, 0, s>>>(count, ne, sigma, cbits, k_in, v_in, sc);
      cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s);
      PP_TRY(pp_dev_alloc(&tmp, tb, s));
      PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s));
      pp_dev_free(k_in, s); pp_dev_free(k_out, s);
    } else {
      uint64_t *k_in, *k_out;
      PP_TRY(pp_dev_alloc(&k_in, ne, s)); PP_TRY(pp_dev_alloc(&k_out, ne, s));
      k_stats_keys<uint64_t><<<std::min(pp_div_up(ne, kBlock), 1184), kBlock, 0, s>>>(count, ne, sigma, cbits, k_in, v_in, sc);
      cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s);
      PP_TRY(pp_dev_alloc(&tmp, tb, s));
      PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, sorted_elem, ne, 0, cbits + wbits, s));
      pp_dev_free(k_in, s); pp_dev_free(k_out, s);
    }
    pp_dev_free(tmp, s); pp_dev_free(v_in, s);
  }
  if (!split)
    k_rows_widths<<<std::min(pp_div_up((long)nchunks * 32, kBlock), 2368), kBlock, 0, s>>>(
        sorted_elem, count, ne, nchunks, L.row_to_element, L.element_to_row, L.row_ppe, width, &sc->cw_sum, &sc->inv);
  pp_dev_free(sorted_elem, s);
  k_chunk_sizes<<<pp_div_up(nchunks + 1, kBlock), kBlock, 0, s>>>(width, nchunks, sc, cfg.shuffle_padding,
                                                                 cfg.padding_strat, V, C, sizes);
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, tb, sizes, pref, Int2Sum(), make_int2(0, 0), nchunks + 1, s);
    char* tmp;
    PP_TRY(pp_dev_alloc(&tmp, tb, s));
    PP_CUDA(cub::DeviceScan::ExclusiveScan(tmp, tb, sizes, pref, Int2Sum(), make_int2(0, 0), nchunks + 1, s));
    pp_dev_free(tmp, s);
  }
  // slices <= one per non-empty chunk + total width / V; total width <= particles * (1 + pad) + chunks
  const double padf = cfg.shuffle_padding > 0 ? cfg.shuffle_padding : 0.0;
  const long nslices_bound = nchunks + (long)((((long)cap + n_new) * (1.0 + padf) + nchunks) / V) + 2;
  PP_TRY(pp_dev_alloc(&L.slice_to_chunk, nslices_bound + 1, s));
  PP_TRY(pp_dev_alloc(&L.offsets, nslices_bound + 2, s));
  PP_TRY(pp_dev_alloc(&L.chunk_start, nchunks + 1, s));
  k_fill_layout<<<pp_div_up(nchunks + 1, kBlock), kBlock, 0, s>>>(pref, width, nchunks, V, C, L.slice_to_chunk,
                                                                 L.offsets, L.chunk_start, sc);
  // chunks in ascending order of their first row's element: the order the gather takes them in
  int* order = nullptr;
  // split layout: rows in front of ne - m are empty for certain, their chunks need no ordering
  const int order_lo = split ? (ne - m_split) / 32 : 0;
  if (g_rebuild_chunk_order) PP_TRY(chunk_order_build(L.row_to_element, width, nchunks, s, &order, order_lo));
  pp_dev_free(width, s); pp_dev_free(sizes, s); pp_dev_free(pref, s);
  delete t_build;                                // SCS_rebuild.h:196-265
  // ---- the one host read
  FastScal h;
  {
    PP_TIME_KIND(s, ps->cfg.kind, "rebuild host read");   // the GPU idles from the copy until the host is back
    static_assert(sizeof(FastScal) <= 256, "FastScal must fit the pinned scratch");
    FastScal* hp = static_cast<FastScal*>(pp_pinned_scratch());
    PP_CUDA(cudaMemcpyAsync(hp ? hp : &h, sc, sizeof(FastScal), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    if (hp) h = *hp;
  }
  auto drop = [&]() {
    free_layout(L, s);
    pp_dev_free(count, s); pp_dev_free(rank, s); pp_dev_free(kept, s); pp_dev_free(order, s); pp_dev_free(sc, s);
    pp_dev_free(rank_new, s);
  };
  if (h.bad) {   // SCS_rebuild.h:147-151 (the reference exits the process)
    drop();
    pp_set_error("there are new particles being added that are marked as inactive (element id -1)");
    return PP_ERR_INVALID;
  }
  int mbits = 1;
  while (mbits < 31 && (1ll << mbits) <= (long long)h.maxcount) ++mbits;
  ps->ppe_bits_hint = mbits + 1;                 // head room: a row may double before the guess fails
  ps->nnz_hint = h.nnz;
  if (h.nnz < C || mbits > cbits || h.active == 0 || h.nslices > nslices_bound || (split && h.nnz > m_split)) {
    drop();                                      // chunk height / key width guessed wrong, or nothing left
    return PP_OK;
  }
  PPTimeScope* t_fin = new PPTimeScope(s, (std::string(pp_kind_name(ps->cfg.kind)) + " layout finish").c_str());
  L.nslices = h.nslices; L.capacity = h.capacity; L.nonempty_chunks = h.cw_cnt;
  const int ntiles = (L.capacity + 31) / 32;
  PP_TRY(pp_dev_alloc(&L.tile_slice, ntiles + 1, s));
  if (ntiles > 0)
    k_tile_slice<<<pp_div_up(ntiles, kBlock), kBlock, 0, s>>>(L.offsets, L.nslices, ntiles, L.tile_slice);
  L.mask_words = ntiles + 1;
  PP_TRY(pp_dev_alloc(&L.mask, L.mask_words, s));
  PP_CUDA(cudaMemsetAsync(L.mask + ntiles, 0, sizeof(uint32_t), s));
  PP_TRY(ensure_swap(ps, L.capacity, s));
  delete t_fin;
  std::vector<const void*> old_src(ps->data.begin(), ps->data.end());
  UnitTable ut, un;
  unit_table(ps, old_src.data(), ps->stride, &ps->swap, ps->swap_stride, ut, remap);
  un = ut;
  if (!g_sm_count_scs) {
    int dev = 0;
    PP_CUDA(cudaGetDevice(&dev));
    PP_CUDA(cudaDeviceGetAttribute(&g_sm_count_scs, cudaDevAttrMultiProcessorCount, dev));
  }
  int* slots = nullptr;
  PP_TIME_KIND(s, ps->cfg.kind, "PSToPs");     // the record move (SCS_rebuild.h:268-271), incl. the new particles
  // wide rows: the 8-byte gathers of a particle cost one L1 wavefront each and the source footprint of
  // a chunk outgrows what L2 can keep for its neighbours; the record stage (full sectors both ways)
  // is faster there (measured: 50 M particles at 25 per element, 5.3 ms staged vs 6.3 ms gathered;
  // 10 M at 10 per element, 1.08 ms staged vs 0.98 ms gathered)
  const int nfull = h.cw_cnt > 0 ? h.cw_cnt : 1;              // chunks that hold particles
  const double avg_cols = (double)L.capacity / (32.0 * nfull);
  // (a chunk wider than V columns -- more slices than chunks -- would be gathered by a single block:
  //  the stage's kernels are thread-per-slot and do not care)
  if (avg_cols <= g_gather_max_cols && L.nslices <= nfull) {
    // ---- single-pass gather
    int* src_of;
    PP_TRY(pp_dev_alloc(&src_of, (size_t)L.capacity + 1, s));
    if (cap > 0)
      k_invmap<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(ps->view(), new_element, rank, L.element_to_row,
                                                        L.chunk_start, src_of);
    if (n_new > 0) {
      k_invmap_new_ranked<<<new_grid, kBlock, 0, s>>>(new_particle_elements, rank_new, n_new, L.element_to_row,
                                                      L.chunk_start, src_of, nullptr, n_new_dev);
      unit_table(ps, new_particle_info, new_ld, &ps->swap, ps->swap_stride, un, remap);
    }
    // blocks in flight: their chunks' source sectors (fetched as whole 64-byte DRAM atoms) must fit L2
    const double foot = 32.0 * avg_cols * (ut.nunits * 8) * 2.0;
    long blocks = g_gather_bps > 0 ? (long)g_sm_count_scs * g_gather_bps : (long)(g_gather_l2_bytes / (foot > 1 ? foot : 1));
    blocks = std::max<long>(g_sm_count_scs, std::min<long>(blocks, (long)g_sm_count_scs * 7));
    // with the chunk order, the non-empty chunks are its first `nfull` entries
    const int ngather = order ? nfull : nchunks;
    const int grid = (int)std::min<long>(ngather, blocks);
    k_gather_scs<4><<<grid, 4 * 32, 0, s>>>(L.chunk_start, L.row_ppe, order, ngather, ut, un, src_of, L.mask,
                                            &sc->next_chunk);
    pp_dev_free(src_of, s);
  } else {
    // ---- record stage: pack in source order, unpack in destination order
    PP_TRY(stage_ensure(ps, (size_t)L.capacity * ut.rec_stride, s));
    if (cap > 0)
      PP_TRY(launch_stage_pack(ps->view(), new_element, L.element_to_row, L.chunk_start, C, 0, nullptr, nullptr,
                               rank, ut, ps->stage, s));
    if (n_new > 0) {
      PP_TRY(pp_dev_alloc(&slots, n_new, s));
      k_invmap_new_ranked<<<new_grid, kBlock, 0, s>>>(new_particle_elements, rank_new, n_new, L.element_to_row,
                                                      L.chunk_start, nullptr, slots, n_new_dev);
      unit_table(ps, new_particle_info, new_ld, nullptr, 0, un, remap);
      k_stage_pack_new<<<new_grid, kBlock, 0, s>>>(slots, n_new, un, ps->stage, n_new_dev);
    }
    k_stage_unpack_scs<<<pp_div_up(L.capacity, kBlock), kBlock, 0, s>>>(layout_view(L, ne), L.row_ppe, ut,
                                                                        ps->stage, L.mask);
  }
  pp_dev_free(slots, s); pp_dev_free(rank_new, s);
  PP_KERNEL_CHECK();
  adopt_layout(ps, L, s);
  std::swap(ps->data, ps->swap);
  std::swap(ps->stride, ps->swap_stride);
  std::swap(ps->data_alloc, ps->swap_alloc);
  if (cfg.always_realloc) {
    for (void* p : ps->swap) pp_dev_free((char*)p, s);
    ps->swap.clear();
    ps->swap_stride = 0;
  }
  ps->nptcls = h.active;
  ps->first_chunk = (cfg.sigma >= ne && h.cw_cnt <= nchunks) ? nchunks - h.cw_cnt : 0;
  pp_dev_free(order, s);
  pp_dev_free(count, s); pp_dev_free(rank, s); pp_dev_free(kept, s); pp_dev_free(sc, s);
  done = true;
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// rebuild (SCS_rebuild.h:123-314; CSR_rebuild.hpp:18-118; dps_rebuild.hpp)
// ------------------------------------------------------------------------------------------
extern "C" pp_status pp_ps_rebuild(pp_ps* ps, const int32_t* new_element, int32_t n_new,
                                   const int32_t* new_particle_elements,
                                   const void* const* new_particle_info, pp_stream stream_) {
  return pp_ps_rebuild_ex(ps, new_element, n_new, nullptr, n_new, new_particle_elements, new_particle_info,
                          stream_);
}

// n_new_dev != NULL: the number of particles being added is *n_new_dev (device memory), n_new is an
// upper bound; new_ld: row length of the new_particle_info arrays ([ncomp][new_ld]).
pp_status pp_ps_rebuild_ex(pp_ps* ps, const int32_t* new_element, int32_t n_new, const int32_t* n_new_dev,
                           int64_t new_ld, const int32_t* new_particle_elements,
                           const void* const* new_particle_info_in, pp_stream stream_) {
  PP_REQUIRE(ps && (new_element || ps->capacity == 0), "null argument");
  PP_REQUIRE(n_new >= 0, "negative number of new particles");
  PP_REQUIRE(n_new == 0 || (new_particle_elements && new_particle_info_in),
             "new particles need their elements and member data");
  cudaStream_t s = (cudaStream_t)stream_;
  const int ne = ps->nelems;
  const int kind = ps->cfg.kind;
  PP_TIME_KIND(s, kind, "rebuild");              // SCS_rebuild.h:312, CSR_rebuild.hpp:116
  const void* const* new_particle_info = new_particle_info_in;
  std::vector<int> remap_v;
  remap_v.swap(ps->rebuild_remap);               // one-shot (pp_ps_set_rebuild_remap)
  const int* remap = remap_v.empty() ? nullptr : remap_v.data();
  // ---- Sell-C-sigma with C = 32, sparse rows: device-side layout + single-pass move
  if ((kind == PP_PS_SCS || kind == PP_PS_CABM) && g_staged_rebuild >= 2 && ps->cfg.team_size == 32 &&
      ne >= 32 && (long)ps->nptcls < (long)g_rank_sort_ppe * ne &&
      // a due reshuffle attempt goes through the general path -- unless the number of particles
      // being added is only known to the device (migration over the peer-memory window): the
      // attempt would need it on the host, and a step that receives particles rarely fits in place
      (n_new_dev || !(g_try_shuffling && ps->capacity > 0 && ps->tile_slice && ps->shuffle_skip == 0))) {
    UnitTable probe;
    std::vector<const void*> old_src(ps->data.begin(), ps->data.end());
    if (unit_table(ps, old_src.data(), ps->stride, nullptr, 0, probe)) {
      bool done = false;
      PP_TRY(rebuild_scs_gather(ps, new_element, n_new, n_new_dev, new_ld, new_particle_elements,
                                new_particle_info, remap, s, done));
      if (done) {
        if (g_try_shuffling && ps->shuffle_skip > 0) --ps->shuffle_skip;
        return PP_OK;
      }
    }
  }
  // ---- general path: the host needs the count, and compact [ncomp][n_new] arrays
  std::vector<char*> compacted;
  std::vector<const void*> compact_ptrs;
  if (n_new_dev) {
    int h_n = 0;
    PP_CUDA(cudaMemcpyAsync(&h_n, n_new_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    n_new = h_n < n_new ? h_n : n_new;
  }
  if (n_new > 0 && new_ld != n_new) {
    for (int i = 0; i < ps->nmembers; ++i) {
      const size_t sb = ps->members[i].scalar_bytes;
      char* q;
      PP_TRY(pp_dev_alloc(&q, sb * ps->members[i].ncomp * (size_t)n_new, s));
      PP_CUDA(cudaMemcpy2DAsync(q, sb * n_new, new_particle_info_in[i], sb * new_ld, sb * n_new,
                                ps->members[i].ncomp, cudaMemcpyDeviceToDevice, s));
      compacted.push_back(q);
      compact_ptrs.push_back(q);
    }
    new_particle_info = compact_ptrs.data();
  }
  // member remap on the general path: permute the member arrays (each source feeds at most one
  // destination), zero-fill the rest; the arrays of the particles being added are permuted the same way
  std::vector<const void*> remapped_info;
  if (remap) {
    const int nm = ps->nmembers;
    std::vector<void*> nd(nm, nullptr);
    std::vector<char> used(nm, 0);
    for (int i = 0; i < nm; ++i)
      if (remap[i] >= 0) { nd[i] = ps->data[remap[i]]; used[remap[i]] = 1; }
    for (int i = 0; i < nm; ++i) {
      if (remap[i] >= 0) continue;
      // (a fresh array must be as large as the arrays it joins: they may be reinterpreted later)
      const size_t bytes = (size_t)ps->members[i].scalar_bytes * ps->members[i].ncomp *
                           (size_t)std::max<long>(ps->stride, ps->data_alloc);
      for (int j = 0; j < nm && !nd[i]; ++j)
        if (!used[j] && ps->members[j].scalar_bytes == ps->members[i].scalar_bytes &&
            ps->members[j].ncomp == ps->members[i].ncomp) { nd[i] = ps->data[j]; used[j] = 1; }
      if (!nd[i]) { char* q; PP_TRY(pp_dev_alloc(&q, bytes, s)); nd[i] = q; }
      PP_CUDA(cudaMemsetAsync(nd[i], 0, bytes ? bytes : 1, s));
    }
    for (int j = 0; j < nm; ++j)
      if (!used[j]) pp_dev_free((char*)ps->data[j], s);
    ps->data = nd;
    if (n_new > 0) {
      remapped_info.assign(nm, nullptr);
      for (int i = 0; i < nm; ++i) {
        if (remap[i] >= 0) { remapped_info[i] = new_particle_info[remap[i]]; continue; }
        const size_t bytes = (size_t)ps->members[i].scalar_bytes * ps->members[i].ncomp * (size_t)n_new;
        char* q;
        PP_TRY(pp_dev_alloc(&q, bytes, s));
        PP_CUDA(cudaMemsetAsync(q, 0, bytes ? bytes : 1, s));
        compacted.push_back(q);
        remapped_info[i] = q;
      }
      new_particle_info = remapped_info.data();
    }
  }
  struct FreeCompacted {
    std::vector<char*>& v; cudaStream_t s;
    ~FreeCompacted() { for (char* q : v) pp_dev_free(q, s); }
  } free_compacted{compacted, s};
  int* scal;
  PP_TRY(pp_dev_alloc(&scal, 4, s));
  PP_CUDA(cudaMemsetAsync(scal, 0, 4 * sizeof(int), s));
  MemberTable mt_new;
  {
    std::vector<void*> srcs(ps->nmembers, nullptr);
    for (int i = 0; i < ps->nmembers && n_new > 0; ++i) srcs[i] = (void*)new_particle_info[i];
    PP_TRY(member_table(ps, srcs, ps->data, mt_new));
  }

  if (kind == PP_PS_DPS) {
    // in place: kept particles only change their parent element; holes are refilled
    if (ps->capacity > 0)
      k_dps_update<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(ps->view(), new_element,
                                                                      ps->slot_elem, ps->mask_bits, scal);
    int nkept = 0;
    PP_CUDA(cudaMemcpyAsync(&nkept, scal, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    const long need = (long)nkept + n_new;
    if (need > ps->capacity) {
      // grow: capacity rule of dps.hpp:129-132 applied to the new particle count
      const int new_cap = (int)ceil(ceil(double(need) / 32) * (1 + ps->cfg.extra_padding)) * 32;
      std::vector<void*> nd;
      PP_TRY(pp_ps_alloc_members(ps, nd, new_cap, s));
      for (int i = 0; i < ps->nmembers; ++i) {
        const int sb = ps->members[i].scalar_bytes;
        for (int c = 0; c < ps->members[i].ncomp; ++c)
          PP_CUDA(cudaMemcpyAsync((char*)nd[i] + (size_t)c * new_cap * sb,
                                  (char*)ps->data[i] + (size_t)c * ps->stride * sb,
                                  (size_t)ps->capacity * sb, cudaMemcpyDeviceToDevice, s));
        pp_dev_free((char*)ps->data[i], s);
      }
      ps->data = nd;
      int* nse;
      uint32_t* nm;
      const long nw = (new_cap + 31) / 32 + 1;
      PP_TRY(pp_dev_alloc(&nse, new_cap + 1, s));
      PP_TRY(pp_dev_alloc(&nm, nw, s));
      PP_CUDA(cudaMemsetAsync(nse, 0, sizeof(int) * (new_cap + 1), s));
      PP_CUDA(cudaMemsetAsync(nm, 0, sizeof(uint32_t) * nw, s));
      PP_CUDA(cudaMemcpyAsync(nse, ps->slot_elem, sizeof(int) * ps->capacity, cudaMemcpyDeviceToDevice, s));
      PP_CUDA(cudaMemcpyAsync(nm, ps->mask_bits, sizeof(uint32_t) * ((ps->capacity + 31) / 32),
                              cudaMemcpyDeviceToDevice, s));
      pp_dev_free(ps->slot_elem, s); pp_dev_free(ps->mask_bits, s);
      ps->slot_elem = nse; ps->mask_bits = nm; ps->mask_words_alloc = nw;
      ps->capacity = new_cap; ps->stride = new_cap;
    }
    if (n_new > 0) {
      const long nwords = (ps->capacity + 31) / 32;
      int *cnt, *off, *slots;
      PP_TRY(pp_dev_alloc(&cnt, nwords + 1, s));
      PP_TRY(pp_dev_alloc(&off, nwords + 1, s));
      PP_TRY(pp_dev_alloc(&slots, n_new, s));
      k_dps_hole_count<<<pp_div_up(nwords + 1, kBlock), kBlock, 0, s>>>(ps->mask_bits, nwords, ps->capacity, cnt);
      PP_TRY(scan_exclusive(cnt, off, (int)nwords + 1, s));
      k_dps_fill_holes<<<pp_div_up(nwords, kBlock), kBlock, 0, s>>>(ps->mask_bits, nwords, ps->capacity, off,
                                                                    new_particle_elements, n_new,
                                                                    ps->slot_elem, slots);
      for (int i = 0; i < ps->nmembers; ++i) mt_new.dst[i] = (char*)ps->data[i];
      k_place_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(slots, n_new, mt_new, ps->stride);
      pp_dev_free(cnt, s); pp_dev_free(off, s); pp_dev_free(slots, s);
    }
    ps->nptcls = (int)need;
    PP_KERNEL_CHECK();
    pp_dev_free(scal, s);
    return PP_OK;
  }

  // ---- element-sorted kinds: histogram of destinations (countNewParticles / rebuild_count)
  int* count;
  PP_TRY(pp_dev_alloc(&count, ne + 1, s));
  PP_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ne + 1), s));
  // crowded elements: ranks and counts from a sort instead of per-element atomics (see k_rank_keys)
  int* rank = nullptr;
  int* kept = nullptr;     // kept particles per element (only with ranks and new particles)
  const bool crowded = g_staged_rebuild && kind != PP_PS_DPS && ps->capacity > 0 &&
                       (long)ps->nptcls >= (long)g_rank_sort_ppe * ne;
  if (crowded) {
    const int cap = ps->capacity;
    unsigned *k_in, *k_out;
    int *v_in, *v_out, *first;
    PP_TRY(pp_dev_alloc(&k_in, cap, s)); PP_TRY(pp_dev_alloc(&k_out, cap, s));
    PP_TRY(pp_dev_alloc(&v_in, cap, s)); PP_TRY(pp_dev_alloc(&v_out, cap, s));
    PP_TRY(pp_dev_alloc(&first, ne + 1, s));
    PP_TRY(pp_dev_alloc(&rank, cap, s));
    k_rank_keys<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(ps->view(), new_element, ne, k_in, v_in);
    int bits = 1;
    while ((1u << bits) <= (unsigned)ne && bits < 32) ++bits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, cap, 0, bits, s);
    char* tmp;
    PP_TRY(pp_dev_alloc(&tmp, tb, s));
    PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, v_out, cap, 0, bits, s));
    k_rank_bounds<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(k_out, cap, ne, first, count);
    k_rank_counts<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(first, count, ne);
    k_rank_scatter<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(k_out, v_out, cap, ne, first, rank);
    if (n_new > 0) {
      PP_TRY(pp_dev_alloc(&kept, ne + 1, s));
      PP_CUDA(cudaMemcpyAsync(kept, count, sizeof(int) * ne, cudaMemcpyDeviceToDevice, s));
    }
    pp_dev_free(tmp, s); pp_dev_free(k_in, s); pp_dev_free(k_out, s); pp_dev_free(v_in, s);
    pp_dev_free(v_out, s); pp_dev_free(first, s);
  } else if (ps->capacity > 0) {
    if (g_staged_rebuild) {
      PP_TRY(pp_dev_alloc(&rank, ps->capacity, s));
      if (n_new > 0) PP_TRY(pp_dev_alloc(&kept, ne + 1, s));
    }
    if (g_hist_block) k_hist_kept_block<<<pp_div_up(ps->capacity, kHistSlots), 256, 0, s>>>(ps->view(), new_element, count, rank);
    else k_hist_kept<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(ps->view(), new_element, count, rank);
    if (kept) PP_CUDA(cudaMemcpyAsync(kept, count, sizeof(int) * ne, cudaMemcpyDeviceToDevice, s));
  }
  if (n_new > 0)
    k_hist_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new, count, scal + 1);
  int* tot_dev;
  PP_TRY(pp_dev_alloc(&tot_dev, ne + 2, s));
  PP_TRY(scan_exclusive(count, tot_dev, ne + 1, s));   // tot_dev[ne] = active particles
  int active = 0, bad = 0;
  PP_CUDA(cudaMemcpyAsync(&active, tot_dev + ne, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaMemcpyAsync(&bad, scal + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  if (bad) {   // SCS_rebuild.h:147-151 (the reference exits the process)
    pp_dev_free(count, s); pp_dev_free(tot_dev, s); pp_dev_free(scal, s); pp_dev_free(rank, s); pp_dev_free(kept, s);
    pp_set_error("there are new particles being added that are marked as inactive (element id -1)");
    return PP_ERR_INVALID;
  }

  if (kind == PP_PS_CSR) {
    // CSR: dense element-major array, capacity = 1.05 * particles (CSR_buildFns.hpp:55-93)
    int new_cap = (int)(active * 1.05);
    if (new_cap < active) new_cap = active;
    const long new_stride = new_cap > 0 ? new_cap : 1;
    std::vector<void*> nd;
    PP_TRY(pp_ps_alloc_members(ps, nd, new_stride, s));
    int* fill;
    PP_TRY(pp_dev_alloc(&fill, ne + 1, s));
    PP_CUDA(cudaMemsetAsync(fill, 0, sizeof(int) * (ne + 1), s));
    std::vector<const void*> old_src(ps->data.begin(), ps->data.end());
    UnitTable ut;
    const bool staged = g_staged_rebuild && active > 0 &&
                        unit_table(ps, old_src.data(), ps->stride, &nd, new_stride, ut);
    if (rank && !staged) {   // ranks feed the staged pack only: fall back to atomics
      pp_dev_free(rank, s); rank = nullptr;
      if (kept) { pp_dev_free(kept, s); kept = nullptr; }
    }
    if (kept) k_fill_from_kept<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(kept, ne, nullptr, fill);
    if (staged) {
      PP_TRY(stage_ensure(ps, (size_t)new_cap * ut.rec_stride, s));
      if (ps->capacity > 0)
        PP_TRY(launch_stage_pack(ps->view(), new_element, nullptr, nullptr, 1, 1, tot_dev, fill, rank, ut,
                                 ps->stage, s));
      if (n_new > 0) {
        int* slots;
        PP_TRY(pp_dev_alloc(&slots, n_new, s));
        k_assign_dense<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new, tot_dev, fill, slots);
        UnitTable un;
        unit_table(ps, new_particle_info, n_new, nullptr, 0, un);
        k_stage_pack_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(slots, n_new, un, ps->stage);
        pp_dev_free(slots, s);
      }
      k_stage_unpack_dense<<<pp_div_up(active, kBlock), kBlock, 0, s>>>(active, ut, ps->stage);
    } else {
      MemberTable mt;
      PP_TRY(member_table(ps, ps->data, nd, mt));
      if (ps->capacity > 0)
        k_move_kept<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(
            ps->view(), new_element, nullptr, nullptr, 1, 1, tot_dev, fill, mt, ps->stride, new_stride,
            nullptr, nullptr);
      if (n_new > 0) {
        int* slots;
        PP_TRY(pp_dev_alloc(&slots, n_new, s));
        k_assign_dense<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new, tot_dev, fill, slots);
        for (int i = 0; i < ps->nmembers; ++i) mt_new.dst[i] = (char*)nd[i];
        k_place_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(slots, n_new, mt_new, new_stride);
        pp_dev_free(slots, s);
      }
    }
    for (void* p : ps->data) pp_dev_free((char*)p, s);
    ps->data = nd;
    ps->stride = new_stride;
    pp_dev_free(ps->mask_bits, s); pp_dev_free(ps->slot_elem, s); pp_dev_free(ps->offsets, s);
    const long nwords = (new_cap + 31) / 32 + 1;
    PP_TRY(pp_dev_alloc(&ps->mask_bits, nwords, s));
    ps->mask_words_alloc = nwords;
    k_mask_first_n2<<<pp_div_up(nwords, kBlock), kBlock, 0, s>>>(ps->mask_bits, nwords, active);
    PP_TRY(pp_dev_alloc(&ps->slot_elem, new_cap + 1, s));
    if (new_cap > 0)
      k_expand_offsets2<<<pp_div_up(new_cap, kBlock), kBlock, 0, s>>>(tot_dev, ne, active, new_cap, ps->slot_elem);
    ps->offsets = tot_dev;
    ps->slot_elem_valid = true;
    ps->capacity = new_cap;
    ps->nptcls = active;
    PP_KERNEL_CHECK();
    pp_dev_free(fill, s); pp_dev_free(count, s); pp_dev_free(scal, s); pp_dev_free(rank, s); pp_dev_free(kept, s);
    return PP_OK;
  }

  // ---- SCS / CabM
  if (active == 0) {   // SCS_rebuild.h:169-181: structure keeps its shape, mask cleared
    PP_CUDA(cudaMemsetAsync(ps->mask_bits, 0, sizeof(uint32_t) * ps->mask_words_alloc, s));
    ps->nptcls = 0;
    pp_dev_free(count, s); pp_dev_free(tot_dev, s); pp_dev_free(scal, s); pp_dev_free(rank, s); pp_dev_free(kept, s);
    return PP_OK;
  }
  // tryShuffling (SCS_rebuild.h:183-189).  An attempt costs a pass over the slots; after a failure
  // the next attempts are spaced out (3, 15, 63, 255 rebuilds), after a success every rebuild tries.
  if (g_try_shuffling && ps->capacity > 0 && ps->tile_slice) {
    if (ps->shuffle_skip > 0) {
      --ps->shuffle_skip;
    } else {
      bool done = false;
      {
        PP_TIME_KIND(s, kind, "shuffle attempt");   // SCS_rebuild.h:190
        PP_TRY(try_reshuffle(ps, new_element, n_new, new_particle_elements, new_particle_info, mt_new, s, done));
      }
      if (done) {
        ps->shuffle_streak = 0;
        ps->nptcls = active;
        pp_dev_free(count, s); pp_dev_free(tot_dev, s); pp_dev_free(scal, s); pp_dev_free(rank, s);
        pp_dev_free(kept, s);
        return PP_OK;
      }
      ps->shuffle_streak = ps->shuffle_streak < 4 ? ps->shuffle_streak + 1 : 4;
      ps->shuffle_skip = (1 << (2 * ps->shuffle_streak)) - 1;      // 3, 15, 63, 255 rebuilds
    }
  }
  ScsLayout L;
  PP_TRY(scs_layout(ps->cfg, ne, count, active, s, L));
  PP_TRY(ensure_swap(ps, L.capacity, s));
  int* row_fill;
  PP_TRY(pp_dev_alloc(&row_fill, L.nrows + 1, s));
  PP_CUDA(cudaMemsetAsync(row_fill, 0, sizeof(int) * (L.nrows + 1), s));
  std::vector<const void*> old_src(ps->data.begin(), ps->data.end());
  UnitTable ut;
  const bool staged = g_staged_rebuild && L.C == 32 &&
                      unit_table(ps, old_src.data(), ps->stride, &ps->swap, ps->swap_stride, ut);
  if (rank && !staged) {   // ranks feed the staged pack only: fall back to atomics
    pp_dev_free(rank, s); rank = nullptr;
    if (kept) { pp_dev_free(kept, s); kept = nullptr; }
  }
  if (kept) k_fill_from_kept<<<pp_div_up(ne, kBlock), kBlock, 0, s>>>(kept, ne, L.element_to_row, row_fill);
  if (staged) {
    PP_TRY(stage_ensure(ps, (size_t)L.capacity * ut.rec_stride, s));
    if (ps->capacity > 0)
      PP_TRY(launch_stage_pack(ps->view(), new_element, L.element_to_row, L.chunk_start, L.C, 0, nullptr,
                               row_fill, rank, ut, ps->stage, s));
    if (n_new > 0) {
      int* slots;
      PP_TRY(pp_dev_alloc(&slots, n_new, s));
      k_assign_slots<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new,
                                                                 L.element_to_row, L.chunk_start, L.C,
                                                                 row_fill, slots);
      UnitTable un;
      unit_table(ps, new_particle_info, n_new, nullptr, 0, un);
      k_stage_pack_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(slots, n_new, un, ps->stage);
      pp_dev_free(slots, s);
    }
    // unpack in destination order; the same pass writes the mask of the new structure
    k_stage_unpack_scs<<<pp_div_up(L.capacity, kBlock), kBlock, 0, s>>>(layout_view(L, ne), L.row_ppe, ut,
                                                                        ps->stage, L.mask);
  } else {
    MemberTable mt;
    PP_TRY(member_table(ps, ps->data, ps->swap, mt));
    if (ps->capacity > 0)
      k_move_kept<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(
          ps->view(), new_element, L.element_to_row, L.chunk_start, L.C, 0, nullptr, row_fill, mt,
          ps->stride, ps->swap_stride, nullptr, nullptr);
    if (n_new > 0) {
      int* slots;
      PP_TRY(pp_dev_alloc(&slots, n_new, s));
      k_assign_slots<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(new_particle_elements, n_new,
                                                                 L.element_to_row, L.chunk_start, L.C,
                                                                 row_fill, slots);
      for (int i = 0; i < ps->nmembers; ++i) mt_new.dst[i] = (char*)ps->swap[i];
      k_place_new<<<pp_div_up(n_new, kBlock), kBlock, 0, s>>>(slots, n_new, mt_new, ps->swap_stride);
      pp_dev_free(slots, s);
    }
    // mask of the new structure: the first count(row) columns of every row are occupied
    PsView v = layout_view(L, ne);
    k_scs_mask<<<pp_div_up(L.capacity, kBlock), kBlock, 0, s>>>(v, L.row_ppe, ne, L.mask, L.mask_words);
  }
  PP_KERNEL_CHECK();
  adopt_layout(ps, L, s);
  std::swap(ps->data, ps->swap);
  std::swap(ps->stride, ps->swap_stride);
  std::swap(ps->data_alloc, ps->swap_alloc);
  if (ps->cfg.always_realloc) {
    for (void* p : ps->swap) pp_dev_free((char*)p, s);
    ps->swap.clear();
    ps->swap_stride = 0;
  }
  ps->nptcls = active;
  pp_dev_free(row_fill, s); pp_dev_free(count, s); pp_dev_free(tot_dev, s); pp_dev_free(scal, s);
  pp_dev_free(rank, s); pp_dev_free(kept, s);
  return PP_OK;
}

// one-shot member remap of the next rebuild / migrate: destination member i <- source member
// src_member[i] (-1: zero).  updatePtclPositions (x <- xtgt, xtgt <- 0) folded into the record move.
extern "C" pp_status pp_ps_set_rebuild_remap(pp_ps* ps, const int32_t* src_member, int32_t n) {
  PP_REQUIRE(ps && (n == 0 || (src_member && n == ps->nmembers)), "one entry per member");
  std::vector<int> m(src_member, src_member + n);
  std::vector<char> used((size_t)n, 0);
  bool identity = true;
  for (int i = 0; i < n; ++i) {
    PP_REQUIRE(m[i] >= -1 && m[i] < n, "source member out of range");
    if (m[i] != i) identity = false;
    if (m[i] < 0) continue;
    PP_REQUIRE(!used[m[i]], "a member may feed at most one destination");
    used[m[i]] = 1;
    PP_REQUIRE(ps->members[m[i]].scalar_bytes == ps->members[i].scalar_bytes &&
                   ps->members[m[i]].ncomp == ps->members[i].ncomp,
               "source and destination member must have the same type");
  }
  if (identity) m.clear();
  ps->rebuild_remap = m;
  return PP_OK;
}

extern "C" void pp_ps_set_staged_rebuild(int32_t mode) { g_staged_rebuild = mode < 0 ? 0 : mode > 2 ? 2 : mode; }
extern "C" void pp_ps_set_rebuild_tuning(int32_t gather_blocks_per_sm, int32_t gather_max_cols) {
  g_gather_bps = gather_blocks_per_sm > 0 ? gather_blocks_per_sm : 0;
  if (gather_max_cols >= 0) g_gather_max_cols = gather_max_cols;
}
extern "C" void pp_ps_set_rebuild_chunk_order(int32_t on) { g_rebuild_chunk_order = on ? 1 : 0; }
extern "C" void pp_ps_set_rebuild_split_rows(int32_t on) { g_rebuild_split_rows = on ? 1 : 0; }
extern "C" void pp_ps_set_rebuild_block_histogram(int32_t on) { g_hist_block = on ? 1 : 0; }

extern "C" void pp_ps_set_rank_sort_threshold(int32_t particles_per_element) {
  g_rank_sort_ppe = particles_per_element > 0 ? particles_per_element : 1;
}

extern "C" void pp_ps_set_shuffling(int32_t on) { g_try_shuffling = on ? 1 : 0; }
