// pp_lb.cu -- particle load balancing on the device.
//
// Replaces pumipic::ParticleBalancer (src/pumipic_lb.hpp:32-115):
//   addWeights       pumipic_lb.hpp:133-208 (particle structure), :211-229 (particles per element)
//   balance          pumipic_lb.cpp:478-511  -> pp_host_lb_plan (csrc/pp_host_lb.cpp)
//   selectParticles  pumipic_lb.hpp:231-289 (particle structure), :291-353 (particles per element)
//   repartition / partition  pumipic_lb.hpp:355-381
// The reference keys its device tables by Kokkos::UnorderedMap (sbar id -> graph vertex, sbar id ->
// plan index).  Here the map is resolved once per element at construction (elem_vert[e] = this
// part's vertex of the element's sbar, or -1), so the per-particle work is two dependent 4-byte
// gathers and one counter:
//   * counting: block-private shared-memory histogram over (own sbar vertices + destination
//     ranks), flushed with one global atomic per non-empty bin and block -- the reference issues
//     one fp64 atomic per particle onto a handful of addresses;
//   * selection: a particle's rank inside its sbar comes from a warp-aggregated counter
//     (__match_any_sync), its target part from the cumulative integer quotas of the plan; the
//     counter is read before it is bumped, so once an sbar's quota is filled the rest of the
//     particles only read.  Counts per (sbar, target) are exactly the plan's (capped by the
//     particles available); WHICH particles go is decided by atomic order, as in the reference.
// Not reproduced from the reference's selection (pumipic_lb.hpp:256-263): it advances to an sbar's
// next target only when a particle reads a remaining weight of exactly 0.0, which never happens for
// a fractional weight -- the first target then gets ceil(w) particles and the later ones none.  Here
// every target gets its ceil(w) in turn.
// The neighbour exchanges of the reference (forced weights :176-200, EnGPar's own) collapse into one
// in-place all-reduce of the global weight vector over NCCL.
#include <cub/cub.cuh>

#include <algorithm>
#include <map>

#include "pp_internal.cuh"

struct pp_balancer {
  int nranks, rank, nelems;
  // global sbar table (host), ascending sbar id
  std::vector<int32_t> sbar_ids, parts_off, parts;
  int nverts;                      // graph vertices of all sbars = max over sbars of id + size
  // this part's vertices, ascending: local vertex l <-> global vertex local_vert[l], sbar local_sbar[l]
  std::vector<int32_t> local_vert, local_sbar;
  int nlocal;
  // device
  int* elem_vert;                  // [nelems] local vertex of the element's sbar or -1
  int* elem_owner;                 // [nelems]
  int* local_vert_dev;             // [nlocal]
  int* counts;                     // [nlocal + nranks] particles per own vertex, then per destination rank
  double* weights;                 // [nverts + nranks] global weight vector (own entries filled)
  // plan of this part (host copy + device tables)
  std::vector<int32_t> plan_sbar, plan_part;
  std::vector<double> plan_weight;
  double imbalance[2];
  int* plan_off;                   // [nlocal + 1] first target of each local vertex
  int* plan_tgt;                   // [ntargets] target part
  int* plan_cum;                   // [ntargets] particles to send to targets before this one (same vertex)
  int* plan_total;                 // [nlocal] particles to send from each local vertex
  int* taken;                      // [nlocal] particles that asked so far
  int plan_alloc;
  bool has_plan;
};

namespace {
constexpr int kBlock = 256;
constexpr int kMaxSharedBins = 4096;

__device__ __forceinline__ bool mask_bit(const uint32_t* __restrict__ mask, int slot) {
  return (__ldg(mask + (slot >> 5)) >> (slot & 31)) & 1u;
}

// accumulateWeight (pumipic_lb.hpp:147-168): bin = own vertex of the new element's sbar when the
// particle stays here, nlocal + destination rank when it is already leaving.
__global__ void k_lb_count_ps(int cap, const uint32_t* __restrict__ mask,
                              const int* __restrict__ new_elems, const int* __restrict__ new_procs,
                              const int* __restrict__ elem_vert, int me, int nlocal, int nbins,
                              int nranks, int* __restrict__ counts) {
  extern __shared__ int s_bins[];
  const bool use_shared = nbins <= kMaxSharedBins;
  if (use_shared) {
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
  }
  for (long slot = blockIdx.x * (long)blockDim.x + threadIdx.x; slot < cap;
       slot += (long)gridDim.x * blockDim.x) {
    if (!mask_bit(mask, (int)slot)) continue;
    const int p = new_procs[slot];
    int bin = -1;
    if (p == me) {
      const int e = new_elems[slot];
      if (e != -1) {
        const int v = __ldg(elem_vert + e);
        if (v >= 0) bin = v;
      }
    } else if (p >= 0 && p < nranks) {
      bin = nlocal + p;
    }
    if (bin < 0) continue;
    if (use_shared) atomicAdd(s_bins + bin, 1);
    else atomicAdd(counts + bin, 1);
  }
  if (use_shared) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
      const int c = s_bins[i];
      if (c) atomicAdd(counts + i, c);
    }
  }
}

// accumulateWeight over particles per element (pumipic_lb.hpp:216-223)
__global__ void k_lb_count_array(int nelems, const int* __restrict__ ppe,
                                 const int* __restrict__ elem_vert, int nbins,
                                 int* __restrict__ counts) {
  extern __shared__ int s_bins[];
  const bool use_shared = nbins <= kMaxSharedBins;
  if (use_shared) {
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0;
    __syncthreads();
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nelems; e += gridDim.x * blockDim.x) {
    const int v = __ldg(elem_vert + e);
    const int c = ppe[e];
    if (v < 0 || c <= 0) continue;
    if (use_shared) atomicAdd(s_bins + v, c);
    else atomicAdd(counts + v, c);
  }
  if (use_shared) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
      const int c = s_bins[i];
      if (c) atomicAdd(counts + i, c);
    }
  }
}

// counts -> the global weight vector: own vertices at their global index, the particles this part
// already sends to rank r at nverts + r (summed over parts by the all-reduce: the weight rank r is
// "forced" to take, pumipic_lb.hpp:196-200)
__global__ void k_lb_weights(const int* __restrict__ counts, const int* __restrict__ local_vert,
                             int nlocal, int nranks, int nverts, double* __restrict__ weights) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nlocal) weights[local_vert[i]] = (double)counts[i];
  else if (i < nlocal + nranks) weights[nverts + (i - nlocal)] = (double)counts[i];
}

// target part of the r-th particle leaving vertex v (r < plan_total[v])
__device__ __forceinline__ int lb_target(const int* __restrict__ plan_off,
                                         const int* __restrict__ plan_tgt,
                                         const int* __restrict__ plan_cum, int v, int r) {
  int t = __ldg(plan_off + v);
  const int end = __ldg(plan_off + v + 1);
  while (t + 1 < end && __ldg(plan_cum + t + 1) <= r) ++t;
  return __ldg(plan_tgt + t);
}

// selectNonCoreParticles / selectParticles (pumipic_lb.hpp:246-288): non_core_only restricts the
// pass to particles whose new element belongs to another part.
// A block owns kSelSlots consecutive slots.  Its candidates are counted per sbar vertex in a small
// shared-memory hash table (the atomic's return value is the candidate's rank inside the block), the
// block reserves its share of each vertex's quota with ONE global atomic per vertex, then every
// candidate whose rank lies inside the quota takes the target the plan names for that rank.  (One
// atomic per warp and vertex on the few dozen quota counters serialised in L2: 3.6 ms for 8.75 M
// re-targeted particles over 64 vertices.)
constexpr int kSelPer = 16, kSelSlots = kBlock * kSelPer, kSelTable = 128;
__global__ void __launch_bounds__(kBlock) k_lb_select_ps(
    int cap, const uint32_t* __restrict__ mask, const int* __restrict__ new_elems, int* __restrict__ new_procs,
    const int* __restrict__ elem_vert, const int* __restrict__ elem_owner, int me, int non_core_only,
    const int* __restrict__ plan_off, const int* __restrict__ plan_tgt, const int* __restrict__ plan_cum,
    const int* __restrict__ plan_total, int* __restrict__ taken) {
  __shared__ int s_key[kSelTable], s_cnt[kSelTable], s_base[kSelTable], s_tot[kSelTable];
  for (int i = threadIdx.x; i < kSelTable; i += kBlock) { s_key[i] = -1; s_cnt[i] = 0; }
  __syncthreads();
  int vv[kSelPer], ent[kSelPer], pos[kSelPer];
  const long s0 = (long)blockIdx.x * kSelSlots + threadIdx.x;
#pragma unroll
  for (int j = 0; j < kSelPer; ++j) {
    const long slot = s0 + (long)j * kBlock;
    int v = -1;
    if (slot < cap && mask_bit(mask, (int)slot) && new_procs[slot] == me) {
      const int e = new_elems[slot];
      if (e != -1 && !(non_core_only && __ldg(elem_owner + e) == me)) {
        v = __ldg(elem_vert + e);
        if (v >= 0) {
          const int total = __ldg(plan_total + v);
          // the counter only grows: a filled quota stays filled
          if (total <= 0 || *(volatile int*)(taken + v) >= total) v = -1;
        }
      }
    }
    vv[j] = v; ent[j] = -1; pos[j] = 0;
    if (v >= 0) {
      unsigned h = ((unsigned)v * 2654435761u) >> 25;          // 7 bits
      for (int probe = 0; probe < kSelTable; ++probe) {
        const int k = atomicCAS(&s_key[h], -1, v);
        if (k == -1 || k == v) { ent[j] = (int)h; break; }
        h = (h + 1) & (kSelTable - 1);
      }
      if (ent[j] >= 0) pos[j] = atomicAdd(&s_cnt[ent[j]], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSelTable; i += kBlock) {
    const int v = s_key[i];
    if (v < 0) continue;
    const int total = __ldg(plan_total + v);
    s_tot[i] = total;
    s_base[i] = (*(volatile int*)(taken + v) >= total) ? total : atomicAdd(taken + v, s_cnt[i]);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSelPer; ++j) {
    const int v = vv[j];
    if (v < 0) continue;
    const long slot = s0 + (long)j * kBlock;
    int r, total;
    if (ent[j] >= 0) {
      r = s_base[ent[j]] + pos[j];
      total = s_tot[ent[j]];
    } else {                                   // more than kSelTable vertices in one block: on its own
      total = __ldg(plan_total + v);
      r = atomicAdd(taken + v, 1);
    }
    if (r < total) new_procs[slot] = lb_target(plan_off, plan_tgt, plan_cum, v, r);
  }
}

// selectParticles over particles per element (pumipic_lb.hpp:316-337): element e owns the entries
// [off[e], off[e] + ppe[e]) of new_procs; it reserves that many ranks of its sbar at once.
__global__ void k_lb_select_array(int nelems, const int* __restrict__ ppe, const int* __restrict__ off,
                                  const int* __restrict__ elem_vert, int me,
                                  const int* __restrict__ plan_off, const int* __restrict__ plan_tgt,
                                  const int* __restrict__ plan_cum, const int* __restrict__ plan_total,
                                  int* __restrict__ taken, int* __restrict__ new_procs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  const int n = ppe[e];
  if (n <= 0) return;
  const int start = off[e];
  const int v = __ldg(elem_vert + e);
  int base = 0, total = 0;
  if (v >= 0) {
    total = __ldg(plan_total + v);
    if (total > 0 && *(volatile int*)(taken + v) < total) base = atomicAdd(taken + v, n);
    else total = 0;
  }
  for (int i = 0; i < n; ++i) {
    const int r = base + i;
    new_procs[start + i] = (r < total) ? lb_target(plan_off, plan_tgt, plan_cum, v, r) : me;
  }
}
}  // namespace

static int lb_grid(long n) {
  // grid-stride kernels: a few blocks per SM of the 148 are enough to keep the shared histograms few
  const long want = (n + kBlock - 1) / kBlock;
  return (int)std::max(1L, std::min(want, 148L * 8));
}

extern "C" pp_status pp_balancer_create(int32_t nranks, int32_t rank, int32_t nsbars,
                                        const int32_t* sbar_ids, const int32_t* parts_off,
                                        const int32_t* parts, int32_t nelems,
                                        const int32_t* elem_sbar, const int32_t* elem_owner,
                                        int32_t memspace, pp_comm* comm, pp_stream stream_,
                                        pp_balancer** out) {
  PP_REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks && nelems >= 0 && nsbars >= 0,
             "bad argument");
  PP_REQUIRE(nsbars == 0 || (sbar_ids && parts_off && parts), "null sbar table");
  PP_REQUIRE(nelems == 0 || (elem_sbar && elem_owner), "null element arrays");
  if (comm) PP_REQUIRE(pp_comm_size(comm) == nranks && pp_comm_rank(comm) == rank,
                       "communicator does not match nranks / rank");
  pp_runtime_init();
  cudaStream_t s = (cudaStream_t)stream_;
  // table keyed by id (sorted); entries known to several ranks must agree
  std::map<int32_t, std::vector<int32_t>> table;
  int32_t nverts = 0;
  for (int i = 0; i < nsbars; ++i) {
    std::vector<int32_t> p(parts + parts_off[i], parts + parts_off[i + 1]);
    PP_REQUIRE(!p.empty() && std::is_sorted(p.begin(), p.end()), "sbar parts must be sorted");
    PP_REQUIRE(p.front() >= 0 && p.back() < nranks && sbar_ids[i] >= 0, "sbar part out of range");
    table[sbar_ids[i]] = p;
    nverts = std::max(nverts, sbar_ids[i] + (int32_t)p.size());
  }
  if (comm && nranks > 1) {
    // merge the tables of all ranks: vertex -> (sbar id, part), MAX-reduced over ranks (every
    // rank that knows an sbar writes the same values, the others -1)
    int* d_n;
    PP_TRY(pp_dev_alloc(&d_n, 1, s));
    PP_CUDA(cudaMemcpyAsync(d_n, &nverts, sizeof(int), cudaMemcpyHostToDevice, s));
    PP_TRY(pp_comm_allreduce(comm, d_n, d_n, 1, PP_INT32, PP_MAX, stream_));
    PP_CUDA(cudaMemcpyAsync(&nverts, d_n, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    pp_dev_free(d_n, s);
    std::vector<int32_t> h((size_t)2 * nverts + 1, -1);
    for (const auto& kv : table)
      for (size_t j = 0; j < kv.second.size(); ++j) {
        h[(size_t)kv.first + j] = kv.first;
        h[(size_t)nverts + kv.first + j] = kv.second[j];
      }
    int* d_t;
    PP_TRY(pp_dev_import(&d_t, h.data(), h.size(), PP_HOST, s));
    PP_TRY(pp_comm_allreduce(comm, d_t, d_t, 2 * (int64_t)nverts, PP_INT32, PP_MAX, stream_));
    PP_CUDA(cudaMemcpyAsync(h.data(), d_t, sizeof(int32_t) * 2 * (size_t)nverts, cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    pp_dev_free(d_t, s);
    table.clear();
    for (int v = 0; v < nverts; ++v)
      if (h[(size_t)v] >= 0) table[h[(size_t)v]].push_back(h[(size_t)nverts + v]);
  }
  pp_balancer* b = new pp_balancer();
  b->nranks = nranks; b->rank = rank; b->nelems = nelems; b->nverts = nverts;
  b->elem_vert = b->elem_owner = b->local_vert_dev = b->counts = nullptr;
  b->weights = nullptr;
  b->plan_off = b->plan_tgt = b->plan_cum = b->plan_total = b->taken = nullptr;
  b->plan_alloc = 0; b->has_plan = false;
  b->imbalance[0] = b->imbalance[1] = 1.0;
  b->parts_off.push_back(0);
  std::map<int32_t, int32_t> sbar_to_local;   // sbar_to_vert of the reference (pumipic_lb.cpp:420-421)
  for (const auto& kv : table) {
    b->sbar_ids.push_back(kv.first);
    for (size_t j = 0; j < kv.second.size(); ++j) {
      b->parts.push_back(kv.second[j]);
      if (kv.second[j] == rank) {
        sbar_to_local[kv.first] = (int32_t)b->local_vert.size();
        b->local_vert.push_back(kv.first + (int32_t)j);
        b->local_sbar.push_back(kv.first);
      }
    }
    b->parts_off.push_back((int32_t)b->parts.size());
  }
  b->nlocal = (int)b->local_vert.size();
  // element -> local vertex
  std::vector<int32_t> es((size_t)nelems), ev((size_t)nelems);
  pp_status st = PP_OK;
  if (nelems) {
    if (memspace == PP_HOST) memcpy(es.data(), elem_sbar, sizeof(int32_t) * (size_t)nelems);
    else if (cudaMemcpy(es.data(), elem_sbar, sizeof(int32_t) * (size_t)nelems, cudaMemcpyDeviceToHost) !=
             cudaSuccess) {
      pp_set_error("pp_balancer_create: cannot read elem_sbar");
      st = PP_ERR_CUDA;
    }
  }
  for (int e = 0; e < nelems && st == PP_OK; ++e) {
    auto it = sbar_to_local.find(es[(size_t)e]);
    ev[(size_t)e] = it == sbar_to_local.end() ? -1 : it->second;
  }
  auto fail = [&](pp_status code) { pp_balancer_destroy(b); return code; };
  if (st != PP_OK) return fail(st);
  if ((st = pp_dev_import(&b->elem_vert, ev.data(), (size_t)nelems, PP_HOST, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_import(&b->elem_owner, elem_owner, (size_t)nelems, memspace, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_import(&b->local_vert_dev, b->local_vert.data(), (size_t)b->nlocal, PP_HOST, s)) != PP_OK)
    return fail(st);
  if ((st = pp_dev_alloc(&b->counts, (size_t)b->nlocal + nranks, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_alloc(&b->weights, (size_t)nverts + nranks, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_alloc(&b->plan_off, (size_t)b->nlocal + 1, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_alloc(&b->plan_total, (size_t)b->nlocal, s)) != PP_OK) return fail(st);
  if ((st = pp_dev_alloc(&b->taken, (size_t)b->nlocal, s)) != PP_OK) return fail(st);
  if (cudaMemsetAsync(b->weights, 0, sizeof(double) * ((size_t)nverts + nranks), s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    pp_set_error("pp_balancer_create: device set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(PP_ERR_CUDA);
  }
  *out = b;
  return PP_OK;
}

extern "C" pp_status pp_balancer_destroy(pp_balancer* b) {
  if (!b) return PP_OK;
  cudaFree(b->elem_vert); cudaFree(b->elem_owner); cudaFree(b->local_vert_dev); cudaFree(b->counts);
  cudaFree(b->weights); cudaFree(b->plan_off); cudaFree(b->plan_tgt); cudaFree(b->plan_cum);
  cudaFree(b->plan_total); cudaFree(b->taken);
  delete b;
  return PP_OK;
}

extern "C" pp_status pp_balancer_info(const pp_balancer* b, int32_t* nverts, int32_t* nlocal,
                                      const int32_t** local_verts, const int32_t** local_sbars) {
  PP_REQUIRE(b, "null argument");
  if (nverts) *nverts = b->nverts;
  if (nlocal) *nlocal = b->nlocal;
  if (local_verts) *local_verts = b->local_vert.data();
  if (local_sbars) *local_sbars = b->local_sbar.data();
  return PP_OK;
}

extern "C" pp_status pp_balancer_weights(pp_balancer* b, double** weights_dev, int64_t* n) {
  PP_REQUIRE(b && weights_dev, "null argument");
  *weights_dev = b->weights;
  if (n) *n = (int64_t)b->nverts + b->nranks;
  return PP_OK;
}

static pp_status lb_publish_counts(pp_balancer* b, cudaStream_t s) {
  PP_CUDA(cudaMemsetAsync(b->weights, 0, sizeof(double) * ((size_t)b->nverts + b->nranks), s));
  const int n = b->nlocal + b->nranks;
  k_lb_weights<<<pp_div_up(n, kBlock), kBlock, 0, s>>>(b->counts, b->local_vert_dev, b->nlocal,
                                                        b->nranks, b->nverts, b->weights);
  PP_KERNEL_CHECK();
  b->has_plan = false;
  return PP_OK;
}

extern "C" pp_status pp_balancer_add_weights_ps(pp_balancer* b, pp_ps* ps, const int32_t* new_elems,
                                                const int32_t* new_procs, pp_stream stream_) {
  // an empty structure (capacity 0) has no slot arrays: a rank without particles still takes part
  PP_REQUIRE(b && ps && ((new_elems && new_procs) || ps->capacity == 0), "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  const int nbins = b->nlocal + b->nranks;
  PP_CUDA(cudaMemsetAsync(b->counts, 0, sizeof(int) * (size_t)nbins, s));
  if (ps->capacity > 0) {
    const size_t shm = nbins <= kMaxSharedBins ? sizeof(int) * (size_t)nbins : 0;
    k_lb_count_ps<<<lb_grid(ps->capacity), kBlock, shm, s>>>(ps->capacity, ps->mask_bits, new_elems,
                                                             new_procs, b->elem_vert, b->rank, b->nlocal,
                                                             nbins, b->nranks, b->counts);
    PP_KERNEL_CHECK();
  }
  return lb_publish_counts(b, s);
}

extern "C" pp_status pp_balancer_add_weights_array(pp_balancer* b, const int32_t* ptcls_per_elem,
                                                   pp_stream stream_) {
  PP_REQUIRE(b && (ptcls_per_elem || b->nelems == 0), "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  const int nbins = b->nlocal + b->nranks;
  PP_CUDA(cudaMemsetAsync(b->counts, 0, sizeof(int) * (size_t)nbins, s));
  if (b->nelems > 0) {
    const size_t shm = nbins <= kMaxSharedBins ? sizeof(int) * (size_t)nbins : 0;
    k_lb_count_array<<<lb_grid(b->nelems), kBlock, shm, s>>>(b->nelems, ptcls_per_elem, b->elem_vert,
                                                             nbins, b->counts);
    PP_KERNEL_CHECK();
  }
  return lb_publish_counts(b, s);
}

extern "C" pp_status pp_balancer_balance(pp_balancer* b, pp_comm* comm, double tol, double step_factor,
                                         pp_stream stream_) {
  PP_REQUIRE(b, "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  const int nw = b->nverts + b->nranks;
  b->plan_sbar.clear(); b->plan_part.clear(); b->plan_weight.clear();
  b->imbalance[0] = b->imbalance[1] = 1.0;
  std::vector<int32_t> off((size_t)b->nlocal + 1, 0), total((size_t)b->nlocal, 0), tgt, cum;
  if (b->nranks > 1) {   // one rank: the empty plan (pumipic_lb.cpp:479-482)
    if (comm && pp_comm_size(comm) > 1)
      PP_TRY(pp_comm_allreduce(comm, b->weights, b->weights, nw, PP_FLOAT64, PP_SUM, stream_));
    std::vector<double> w((size_t)nw);
    PP_CUDA(cudaMemcpyAsync(w.data(), b->weights, sizeof(double) * (size_t)nw, cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    int32_t ns = 0, *sv = nullptr, *sp = nullptr;
    double* sw = nullptr;
    PP_TRY(pp_host_lb_plan(b->nranks, (int32_t)b->sbar_ids.size(), b->sbar_ids.data(),
                           b->parts_off.data(), b->parts.data(), b->nverts, w.data(),
                           w.data() + b->nverts, tol, step_factor, &ns, &sv, &sp, &sw, b->imbalance));
    // keep this part's sends, grouped by local vertex in ascending (vertex, target) order
    for (int l = 0; l < b->nlocal; ++l) {
      int acc = 0;
      for (int i = 0; i < ns; ++i) {
        if (sv[i] != b->local_vert[(size_t)l]) continue;
        const int n = (int)ceil(sw[i] - 1e-9);   // the reference sends while weight > 0 remains
        if (n <= 0) continue;
        b->plan_sbar.push_back(b->local_sbar[(size_t)l]);
        b->plan_part.push_back(sp[i]);
        b->plan_weight.push_back(sw[i]);
        tgt.push_back(sp[i]);
        cum.push_back(acc);
        acc += n;
      }
      total[(size_t)l] = acc;
      off[(size_t)l + 1] = (int32_t)tgt.size();
    }
    free(sv); free(sp); free(sw);
  }
  const int nt = (int)tgt.size();
  if (nt > b->plan_alloc) {
    pp_dev_free(b->plan_tgt, s); pp_dev_free(b->plan_cum, s);
    b->plan_tgt = b->plan_cum = nullptr; b->plan_alloc = 0;
    PP_TRY(pp_dev_alloc(&b->plan_tgt, (size_t)nt, s));
    PP_TRY(pp_dev_alloc(&b->plan_cum, (size_t)nt, s));
    b->plan_alloc = nt;
  }
  if (nt) {
    PP_CUDA(cudaMemcpyAsync(b->plan_tgt, tgt.data(), sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, s));
    PP_CUDA(cudaMemcpyAsync(b->plan_cum, cum.data(), sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, s));
  }
  PP_CUDA(cudaMemcpyAsync(b->plan_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, s));
  if (b->nlocal) {
    PP_CUDA(cudaMemcpyAsync(b->plan_total, total.data(), sizeof(int) * total.size(), cudaMemcpyHostToDevice, s));
    PP_CUDA(cudaMemsetAsync(b->taken, 0, sizeof(int) * (size_t)b->nlocal, s));
  }
  PP_CUDA(cudaStreamSynchronize(s));   // the host vectors above go out of scope
  b->has_plan = true;
  return PP_OK;
}

extern "C" pp_status pp_balancer_plan(const pp_balancer* b, int32_t* nsends, const int32_t** sbar,
                                      const int32_t** part, const double** weight,
                                      double imbalance[2]) {
  PP_REQUIRE(b && nsends, "null argument");
  PP_REQUIRE(b->has_plan, "no plan: call pp_balancer_balance first");
  *nsends = (int32_t)b->plan_sbar.size();
  if (sbar) *sbar = b->plan_sbar.data();
  if (part) *part = b->plan_part.data();
  if (weight) *weight = b->plan_weight.data();
  if (imbalance) { imbalance[0] = b->imbalance[0]; imbalance[1] = b->imbalance[1]; }
  return PP_OK;
}

extern "C" pp_status pp_balancer_select_ps(pp_balancer* b, pp_ps* ps, const int32_t* new_elems,
                                           int32_t* new_procs, pp_stream stream_) {
  PP_REQUIRE(b && ps && ((new_elems && new_procs) || ps->capacity == 0), "null argument");
  PP_REQUIRE(b->has_plan, "no plan: call pp_balancer_balance first");
  if (b->nranks == 1 || b->plan_sbar.empty() || ps->capacity <= 0) return PP_OK;
  cudaStream_t s = (cudaStream_t)stream_;
  for (int non_core_only = 1; non_core_only >= 0; --non_core_only) {
    k_lb_select_ps<<<pp_div_up(ps->capacity, kSelSlots), kBlock, 0, s>>>(
        ps->capacity, ps->mask_bits, new_elems, new_procs, b->elem_vert, b->elem_owner, b->rank,
        non_core_only, b->plan_off, b->plan_tgt, b->plan_cum, b->plan_total, b->taken);
    PP_KERNEL_CHECK();
  }
  return PP_OK;
}

extern "C" pp_status pp_balancer_select_array(pp_balancer* b, const int32_t* ptcls_per_elem,
                                              int64_t nptcls, int32_t* new_procs, pp_stream stream_) {
  PP_REQUIRE(b && (ptcls_per_elem || b->nelems == 0) && (new_procs || nptcls == 0), "null argument");
  PP_REQUIRE(b->has_plan, "no plan: call pp_balancer_balance first");
  if (b->nelems == 0 || nptcls == 0) return PP_OK;
  cudaStream_t s = (cudaStream_t)stream_;
  int* off;
  PP_TRY(pp_dev_alloc(&off, (size_t)b->nelems, s));
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, ptcls_per_elem, off, b->nelems, s);
  char* tmp;
  PP_TRY(pp_dev_alloc(&tmp, tb, s));
  PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, ptcls_per_elem, off, b->nelems, s));
  k_lb_select_array<<<pp_div_up(b->nelems, kBlock), kBlock, 0, s>>>(
      b->nelems, ptcls_per_elem, off, b->elem_vert, b->rank, b->plan_off, b->plan_tgt, b->plan_cum,
      b->plan_total, b->taken, new_procs);
  pp_dev_free(tmp, s);
  pp_dev_free(off, s);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_balancer_repartition(pp_balancer* b, pp_comm* comm, pp_ps* ps, double tol,
                                             const int32_t* new_elems, int32_t* new_procs,
                                             double step_factor, pp_stream stream) {
  PP_REQUIRE(b, "null argument");
  if (b->nranks == 1) return PP_OK;   // pumipic_lb.hpp:360-361
  PP_TRY(pp_balancer_add_weights_ps(b, ps, new_elems, new_procs, stream));
  PP_TRY(pp_balancer_balance(b, comm, tol, step_factor, stream));
  return pp_balancer_select_ps(b, ps, new_elems, new_procs, stream);
}

extern "C" pp_status pp_balancer_partition(pp_balancer* b, pp_comm* comm, const int32_t* ptcls_per_elem,
                                           int64_t nptcls, double tol, double step_factor,
                                           int32_t* new_procs, pp_stream stream) {
  PP_REQUIRE(b, "null argument");
  PP_TRY(pp_balancer_add_weights_array(b, ptcls_per_elem, stream));
  PP_TRY(pp_balancer_balance(b, comm, tol, step_factor, stream));
  return pp_balancer_select_array(b, ptcls_per_elem, nptcls, new_procs, stream);
}
