// pp_ps.cu -- particle structures (device-resident): construction, accessors, slot geometry.
//
// Replaces particle_structs/src/{scs,csr,dps,cabm}: ParticleStructure<DataTypes> keeps one
// component-major SoA array per member (support/MemberTypeLibraries.h:17,47-88; ppView.h:7-10),
// a particle mask and the slot -> element map that ps::parallel_for walks.  Here the mask is a
// bit per slot (one 4-byte broadcast load per warp instead of 32 byte loads) and all kinds share
// one storage engine.
#include <cub/cub.cuh>

#include "pp_internal.cuh"

namespace {
constexpr int kBlock = 256;

__global__ void k_mask_first_n(uint32_t* mask, long nwords, int n) {
  long w = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  long lo = w * 32;
  uint32_t v = 0;
  if (lo + 32 <= n) v = 0xffffffffu;
  else if (lo < n) v = (1u << (n - lo)) - 1u;
  mask[w] = v;
}

// slot -> element for an element-sorted dense layout: off = exclusive scan of ppe
__global__ void k_expand_offsets(const int* __restrict__ off, int ne, int n, int cap, int* slot_elem) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  if (s >= n) { slot_elem[s] = 0; return; }
  int lo = 0, hi = ne;  // last e with off[e] <= s
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= s) lo = mid; else hi = mid;
  }
  slot_elem[s] = lo;
}

__global__ void k_copy_ints(const int* __restrict__ src, int n, int cap, int* dst, int fill) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  dst[s] = s < n ? src[s] : fill;
}

__global__ void k_csr_slots(const int* __restrict__ elems, int n, const int* __restrict__ off, int* fill,
                            int* slots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = elems[i];
  slots[i] = off[e] + atomicAdd(fill + e, 1);
}

// copy member data [ncomp][np] -> [ncomp][stride] at slots given by `slots` (or identity)
__global__ void k_place_member(const char* __restrict__ src, long np, int ncomp, int sb,
                               const int* __restrict__ slots, char* dst, long stride) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= np * ncomp) return;
  long p = i % np;
  int c = (int)(i / np);
  long slot = slots ? slots[p] : p;
  const char* s = src + (c * np + p) * sb;
  char* d = dst + (c * stride + slot) * sb;
  if (sb == 8) *(double*)d = *(const double*)s;
  else if (sb == 4) *(int*)d = *(const int*)s;
  else for (int b = 0; b < sb; ++b) d[b] = s[b];
}
}  // namespace

PsView pp_ps::view() const {
  PsView v;
  v.kind = cfg.kind;
  v.capacity = capacity;
  v.mask_bits = mask_bits;
  v.slot_elem = slot_elem_valid ? slot_elem : nullptr;
  v.offsets = offsets;
  v.slice_to_chunk = slice_to_chunk;
  v.row_to_element = row_to_element;
  v.tile_slice = tile_slice;
  v.C = C;
  v.nslices = nslices;
  v.chunk_start = chunk_start;
  v.nchunks = (cfg.kind == PP_PS_SCS || cfg.kind == PP_PS_CABM) ? nchunks : 0;
  v.nelems = nelems;
  v.first_chunk = first_chunk;
  v.sliced = sliced;
  return v;
}

extern "C" void pp_ps_config_default(pp_ps_config* c, int32_t kind) {
  if (!c) return;
  c->kind = kind;
  c->team_size = 32;
  c->sigma = 0x7fffffff;
  c->V = 1024;
  c->shuffle_padding = 0.1;   // SellCSigma.h:283-287
  c->extra_padding = 0.05;
  c->minimize_size = 0.8;
  c->padding_strat = PP_PAD_EVENLY;
  c->always_realloc = 0;
}

pp_status pp_ps_alloc_members(pp_ps* ps, std::vector<void*>& arrs, long stride, cudaStream_t s) {
  arrs.assign(ps->nmembers, nullptr);
  for (int i = 0; i < ps->nmembers; ++i) {
    size_t bytes = (size_t)ps->members[i].scalar_bytes * ps->members[i].ncomp * (size_t)stride;
    char* p;
    PP_TRY(pp_dev_alloc(&p, bytes, s));
    PP_CUDA(cudaMemsetAsync(p, 0, bytes ? bytes : 1, s));
    arrs[i] = p;
  }
  return PP_OK;
}

pp_status pp_scs_build(pp_ps* ps, const int* ppe_dev, const int* pelems_dev,
                       const void* const* pinfo, int memspace, cudaStream_t s);

static pp_status build_flat(pp_ps* ps, const int* ppe_dev, const int* pelems_dev,
                            const void* const* pinfo, int memspace, cudaStream_t s) {
  const int np = ps->nptcls, ne = ps->nelems;
  const bool csr = ps->cfg.kind == PP_PS_CSR;
  // dps.hpp:129-132: capacity = ceil(ceil(np/VL)*(1+extra_padding))*VL with VL = 32 on the GPU;
  // CSR_buildFns.hpp:55-93: capacity = np * 1.05
  if (csr) ps->capacity = (int)(np * 1.05);
  else ps->capacity = (int)ceil(ceil(double(np) / 32) * (1 + ps->cfg.extra_padding)) * 32;
  if (ps->capacity < np) ps->capacity = np;
  ps->stride = ps->capacity > 0 ? ps->capacity : 1;
  ps->nrows = ne;
  const long nwords = (ps->capacity + 31) / 32 + 1;
  PP_TRY(pp_dev_alloc(&ps->mask_bits, nwords, s));
  ps->mask_words_alloc = nwords;
  k_mask_first_n<<<pp_div_up(nwords, kBlock), kBlock, 0, s>>>(ps->mask_bits, nwords, np);
  PP_TRY(pp_dev_alloc(&ps->slot_elem, ps->capacity + 1, s));
  PP_TRY(pp_ps_alloc_members(ps, ps->data, ps->stride, s));
  int* off;
  PP_TRY(pp_dev_alloc(&off, ne + 1, s));
  {
    // offsets over elements (CSR.hpp offsets; also the DPS "no particle data" parent fill)
    int* cnt;
    PP_TRY(pp_dev_alloc(&cnt, ne + 1, s));
    PP_CUDA(cudaMemcpyAsync(cnt, ppe_dev, sizeof(int) * ne, cudaMemcpyDeviceToDevice, s));
    PP_CUDA(cudaMemsetAsync(cnt + ne, 0, sizeof(int), s));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, ne + 1, s);
    char* tmp;
    PP_TRY(pp_dev_alloc(&tmp, tb, s));
    PP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, ne + 1, s));
    pp_dev_free(tmp, s);
    pp_dev_free(cnt, s);
  }
  if (ps->capacity > 0) {
    if (!csr && pelems_dev)  // dps fillAoSoA: particle i -> slot i, parent = particle_elements[i]
      k_copy_ints<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(pelems_dev, np, ps->capacity,
                                                                      ps->slot_elem, 0);
    else
      k_expand_offsets<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(off, ne, np, ps->capacity,
                                                                           ps->slot_elem);
  }
  ps->slot_elem_valid = true;
  if (pinfo && np > 0) {
    int* slots = nullptr;
    if (csr) {   // CSR_buildFns.hpp:95-141 initCsrData: particle i -> next free slot of its element
      PP_REQUIRE(pelems_dev, "CSR initial particle data needs particle_elements");
      int* fill;
      PP_TRY(pp_dev_alloc(&fill, ne + 1, s));
      PP_TRY(pp_dev_alloc(&slots, np, s));
      PP_CUDA(cudaMemsetAsync(fill, 0, sizeof(int) * (ne + 1), s));
      k_csr_slots<<<pp_div_up(np, kBlock), kBlock, 0, s>>>(pelems_dev, np, off, fill, slots);
      pp_dev_free(fill, s);
    }
    for (int i = 0; i < ps->nmembers; ++i) {
      const int sb = ps->members[i].scalar_bytes, nc = ps->members[i].ncomp;
      char* src;
      PP_TRY(pp_dev_import(&src, (const char*)pinfo[i], (size_t)sb * nc * np, memspace, s));
      k_place_member<<<pp_div_up((long)np * nc, kBlock), kBlock, 0, s>>>(src, np, nc, sb, slots,
                                                                         (char*)ps->data[i], ps->stride);
      pp_dev_free(src, s);
    }
    pp_dev_free(slots, s);
  }
  if (csr) ps->offsets = off; else pp_dev_free(off, s);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_ps_create(const pp_ps_config* cfg, int32_t nmembers,
                                  const pp_member_desc* members, int32_t ne, int32_t np,
                                  const int32_t* ppe, const int64_t* elem_gids,
                                  const int32_t* particle_elements,
                                  const void* const* particle_info, int32_t memspace,
                                  pp_stream stream_, pp_ps** out) {
  PP_REQUIRE(cfg && members && out && ppe, "null argument");
  PP_REQUIRE(nmembers > 0 && ne > 0 && np >= 0, "bad sizes");
  for (int i = 0; i < nmembers; ++i)
    PP_REQUIRE(members[i].scalar_bytes > 0 && members[i].ncomp > 0, "bad member descriptor");
  pp_runtime_init();
  cudaStream_t s = (cudaStream_t)stream_;
  pp_ps* ps = new pp_ps();
  ps->cfg = *cfg;
  ps->nmembers = nmembers;
  ps->members.assign(members, members + nmembers);
  ps->nelems = ne; ps->nptcls = np; ps->capacity = 0; ps->nrows = 0; ps->stride = 0;
  ps->swap_stride = 0; ps->mask_bits = nullptr; ps->mask_words_alloc = 0;
  ps->slot_elem = nullptr; ps->slot_elem_valid = false; ps->slot_elem_materialized = false;
  ps->chunk_start = nullptr; ps->row_ppe = nullptr;
  ps->C = 1; ps->V = cfg->V; ps->nchunks = 0; ps->nslices = 0;
  ps->offsets = ps->slice_to_chunk = ps->row_to_element = ps->element_to_row = ps->tile_slice = nullptr;
  ps->elem_gids = nullptr; ps->sorted_gid = nullptr; ps->sorted_lid = nullptr;
  ps->stage = nullptr; ps->stage_bytes = 0;
  int* ppe_dev;
  PP_TRY(pp_dev_import(&ppe_dev, ppe, (size_t)ne, memspace, s));
  int* pel_dev = nullptr;
  if (particle_elements) PP_TRY(pp_dev_import(&pel_dev, particle_elements, (size_t)np, memspace, s));
  if (elem_gids) PP_TRY(pp_dev_import(&ps->elem_gids, elem_gids, (size_t)ne, memspace, s));
  pp_status st;
  if (cfg->kind == PP_PS_DPS || cfg->kind == PP_PS_CSR)
    st = build_flat(ps, ppe_dev, pel_dev, particle_info, memspace, s);
  else
    st = pp_scs_build(ps, ppe_dev, pel_dev, particle_info, memspace, s);
  pp_dev_free(ppe_dev, s);
  pp_dev_free(pel_dev, s);
  if (st != PP_OK) { pp_ps_destroy(ps); return st; }
  PP_CUDA(cudaStreamSynchronize(s));
  *out = ps;
  return PP_OK;
}

extern "C" pp_status pp_ps_destroy(pp_ps* ps) {
  if (!ps) return PP_OK;
  for (void* p : ps->data) cudaFree(p);
  for (void* p : ps->swap) cudaFree(p);
  cudaFree(ps->mask_bits); cudaFree(ps->slot_elem); cudaFree(ps->offsets);
  cudaFree(ps->slice_to_chunk); cudaFree(ps->row_to_element); cudaFree(ps->element_to_row);
  cudaFree(ps->tile_slice); cudaFree(ps->elem_gids); cudaFree(ps->chunk_start); cudaFree(ps->row_ppe); cudaFree(ps->sorted_gid); cudaFree(ps->sorted_lid); cudaFree(ps->stage);
  delete ps;
  return PP_OK;
}

extern "C" int32_t pp_ps_nelems(const pp_ps* ps) { return ps ? ps->nelems : -1; }
extern "C" int32_t pp_ps_nptcls(const pp_ps* ps) { return ps ? ps->nptcls : -1; }
extern "C" int32_t pp_ps_capacity(const pp_ps* ps) { return ps ? ps->capacity : -1; }
extern "C" int32_t pp_ps_numrows(const pp_ps* ps) { return ps ? ps->nrows : -1; }
extern "C" int32_t pp_ps_kind_of(const pp_ps* ps) { return ps ? ps->cfg.kind : -1; }

extern "C" pp_status pp_ps_member(const pp_ps* ps, int32_t i, void** base, int64_t* stride) {
  PP_REQUIRE(ps && base && stride, "null argument");
  PP_REQUIRE(i >= 0 && i < ps->nmembers, "member index out of range");
  *base = ps->data[i];
  *stride = ps->stride;
  return PP_OK;
}

namespace {
__global__ void k_materialize_slot_elem(PsView v, int* out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  int e;
  pp_slot_lookup(v, s, e);
  out[s] = e;
}
}  // namespace

extern "C" pp_status pp_ps_get_layout(pp_ps* ps, pp_stream stream_, pp_ps_layout* o) {
  PP_REQUIRE(ps && o, "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  if (!ps->slot_elem_valid && !ps->slot_elem_materialized && ps->capacity > 0) {
    if (!ps->slot_elem) PP_TRY(pp_dev_alloc(&ps->slot_elem, ps->capacity + 1, s));
    k_materialize_slot_elem<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(ps->view(), ps->slot_elem);
    PP_KERNEL_CHECK();
    // kernels keep the SCS tile lookup: slot_elem_valid stays false, the copy is for callers
    ps->slot_elem_materialized = true;
  }
  o->kind = ps->cfg.kind; o->C = ps->C; o->V = ps->V; o->nchunks = ps->nchunks;
  o->nslices = ps->nslices; o->nrows = ps->nrows; o->capacity = ps->capacity;
  o->nelems = ps->nelems; o->nptcls = ps->nptcls;
  o->offsets = ps->offsets; o->slice_to_chunk = ps->slice_to_chunk;
  o->row_to_element = ps->row_to_element; o->element_to_row = ps->element_to_row;
  o->mask_bits = ps->mask_bits; o->slot_elem = ps->slot_elem;
  return PP_OK;
}

// ------------------------------------------------------------------------------------------
// getPIDs (particle_structs/src/ps_for.hpp:57-88): slots of the masked particles grouped by
// element + the start of each element's group.  The reference orders a group by atomic arrival;
// here a stable radix sort of (element, slot) makes it ascending in slot, which is one of the
// orders the reference can produce.
// ------------------------------------------------------------------------------------------
namespace {
__global__ void k_pid_keys(PsView v, int nelems, int* __restrict__ keys, int* __restrict__ slots) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  int e;
  const bool m = pp_slot_lookup(v, s, e);
  keys[s] = m ? e : nelems;   // unmasked slots sort behind every element
  slots[s] = s;
}
// offsets[e] = first position of a key >= e in the sorted keys, e = 0..nelems
__global__ void k_pid_offsets(const int* __restrict__ sorted_keys, int n, int nelems,
                              int* __restrict__ offsets) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nelems) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < e) lo = mid + 1; else hi = mid;
  }
  offsets[e] = lo;
}
}  // namespace

extern "C" pp_status pp_ps_get_pids(pp_ps* ps, int32_t* pids, int32_t* offsets, pp_stream stream_) {
  PP_REQUIRE(ps && offsets && (pids || ps->nptcls == 0), "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  const int cap = ps->capacity, ne = ps->nelems;
  if (cap == 0) {
    PP_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int) * ((size_t)ne + 1), s));
    return PP_OK;
  }
  int *keys, *slots, *keys_out, *slots_out;
  PP_TRY(pp_dev_alloc(&keys, (size_t)cap, s));
  PP_TRY(pp_dev_alloc(&slots, (size_t)cap, s));
  PP_TRY(pp_dev_alloc(&keys_out, (size_t)cap, s));
  PP_TRY(pp_dev_alloc(&slots_out, (size_t)cap, s));
  k_pid_keys<<<pp_div_up(cap, kBlock), kBlock, 0, s>>>(ps->view(), ne, keys, slots);
  int bits = 1;
  while (bits < 31 && (1 << bits) <= ne) ++bits;   // keys are 0..ne
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_out, slots, slots_out, cap, 0, bits, s);
  char* tmp;
  PP_TRY(pp_dev_alloc(&tmp, tb, s));
  PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_out, slots, slots_out, cap, 0, bits, s));
  k_pid_offsets<<<pp_div_up(ne + 1, kBlock), kBlock, 0, s>>>(keys_out, cap, ne, offsets);
  if (ps->nptcls > 0)
    PP_CUDA(cudaMemcpyAsync(pids, slots_out, sizeof(int) * (size_t)ps->nptcls, cudaMemcpyDeviceToDevice, s));
  pp_dev_free(tmp, s);
  pp_dev_free(keys, s); pp_dev_free(slots, s); pp_dev_free(keys_out, s); pp_dev_free(slots_out, s);
  PP_KERNEL_CHECK();
  return PP_OK;
}
