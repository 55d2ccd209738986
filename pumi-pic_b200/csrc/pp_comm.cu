// pp_comm.cu -- NCCL transport, particle migration and the ghost-entity field reduction.
//
// Replaces support/ViewComm*.h(pp) (PS_Comm_* MPI wrappers, host-staged unless GPU-aware MPI),
// particle_structs/src/scs/SCS_migrate.h:5-221 (one Isend/Irecv per member type per peer, each
// preceded by a pack kernel + fence) and the full-mesh branch of Mesh::reduceCommArray
// (src/pumipic_comm.cpp:223-247, a host-staged MPI_Allreduce).  Here everything stays in HBM:
// counts travel in one ncclAllGather, every peer gets ONE packed byte buffer inside a
// ncclGroupStart/End (= all-to-all-v over NVLink), and the reduction is an in-place ncclAllReduce.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): in a torch process that is the NCCL torch
// already loaded, otherwise the system library.  Single-GPU use never touches it.
#include <dlfcn.h>

#include <algorithm>
#include <nccl.h>

#include <cub/cub.cuh>

#include "pp_internal.cuh"

namespace {
constexpr int kBlock = 256;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

pp_status load_nccl() {
  if (g_nccl.handle) return PP_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    pp_set_error("cannot load NCCL: %s", dlerror());
    return PP_ERR_NCCL;
  }
#define PP_SYM(field, name)                                        \
  *(void**)(&g_nccl.field) = dlsym(h, name);                       \
  if (!g_nccl.field) { pp_set_error("NCCL symbol %s missing", name); return PP_ERR_NCCL; }
  PP_SYM(GetUniqueId, "ncclGetUniqueId");
  PP_SYM(CommInitRank, "ncclCommInitRank");
  PP_SYM(CommDestroy, "ncclCommDestroy");
  PP_SYM(AllReduce, "ncclAllReduce");
  PP_SYM(AllGather, "ncclAllGather");
  PP_SYM(Send, "ncclSend");
  PP_SYM(Recv, "ncclRecv");
  PP_SYM(GroupStart, "ncclGroupStart");
  PP_SYM(GroupEnd, "ncclGroupEnd");
  PP_SYM(GetErrorString, "ncclGetErrorString");
#undef PP_SYM
  g_nccl.handle = h;
  return PP_OK;
}

#define PP_NCCL(call)                                                                  \
  do {                                                                                 \
    ncclResult_t r__ = (call);                                                         \
    if (r__ != ncclSuccess) {                                                          \
      pp_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                  \
                   g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");          \
      return PP_ERR_NCCL;                                                              \
    }                                                                                  \
  } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------
// Peer-memory window of the migration (NVLink / NVSwitch, no NCCL and no host in the step).
// Every rank owns one cudaMalloc'ed window: a header and, for both parities of the step counter,
// one fixed-size segment per sender.  The windows are mapped into every peer with CUDA IPC once.
// A migration then is: pack kernel stores each leaving particle's record straight into the
// destination rank's segment (P2P stores), a publish kernel writes the per-peer counts and the
// step number into the destinations' headers, a wait kernel spins on this rank's own header until
// every sender has published, an unpack kernel turns the received records into the rebuild's
// new-particle arrays -- and the number of received particles never leaves the device (the rebuild
// takes it from device memory).  Segments are double-buffered by step parity: a sender can only
// reach step e after it has seen every peer's flag of step e-1, which that peer wrote after it
// finished unpacking step e-2, so a segment is never overwritten while it is still being read.
// ------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 64;
struct P2PHeader {
  int counts[2][kMaxRanks];        // [parity][sender]: particles the sender put into its segment
  unsigned flags[2][kMaxRanks];    // [parity][sender]: step in which the sender finished writing
};
struct P2PPeers { char* win[kMaxRanks]; };
struct P2PWindow {
  bool tried = false, ok = false;
  size_t seg_bytes = 24u << 20;    // per (sender, parity); the same on every rank after the set-up
  bool seg_set = false;            // sized by pp_comm_set_p2p_window
  char* local = nullptr;
  P2PPeers peers;
  int* dev = nullptr;              // cursor[R] | overflow | err | n_in | n_total | recv_off[R+1] | recv_cnt[R]
  long long* host_stats = nullptr; // pinned: sent, received, deferred, err
  long long* dev_stats = nullptr;
  cudaEvent_t ev = nullptr;
  unsigned epoch = 0;
};
// Peer-memory window of the comm-array reduction (Mesh::reduceCommArray, all-reduce branch): header,
// input copy A, result B.  Reduce-scatter + all-gather by direct loads / stores over NVLink: rank r
// reduces slice r of everybody's A in ascending rank order (the same bits on every rank and in every
// run) and stores the result into everybody's B.
struct P2PReduceHeader {
  unsigned ready[kMaxRanks];       // [rank]: call number whose input copy that rank has finished
  unsigned done[kMaxRanks];        // [rank]: call number whose slice that rank has stored everywhere
};
struct P2PReduce {
  bool tried = false, ok = false;
  size_t cap_bytes = 0;            // bytes of A (and of B)
  char* local = nullptr;
  P2PPeers peers;
  int* err = nullptr;              // device flag: a peer did not show up
  unsigned epoch = 0;
};
int g_p2p_enable = 1;

struct pp_comm {
  ncclComm_t comm;                 // null: one rank, or a hosted communicator without NCCL
  int nranks, rank;
  pp_host_allgather_fn host_ag = nullptr;   // hosted: the application's all-gather (MPI_Allgather, ...)
  void* host_ctx = nullptr;
  P2PWindow p2p;
  P2PReduce red;
};
#define PP_NEED_NCCL(c)                                                                               \
  PP_REQUIRE((c)->comm, "this communicator has no NCCL transport (pp_comm_create_hosted without NCCL): only " \
                        "the peer-memory paths (migrate, array_reduce, allreduce) are available")

namespace {
// Set-up collective of the communicator: every rank contributes `bytes` bytes of HOST memory, all ranks
// get the nranks blocks in rank order.  Hosted communicators use the application's function, the
// others NCCL on staging buffers.  Only the one-time window set-ups use it, never a step.
pp_status boot_allgather(pp_comm* c, const void* h_send, void* h_recv, size_t bytes, cudaStream_t s) {
  if (c->nranks == 1) { memcpy(h_recv, h_send, bytes); return PP_OK; }
  if (c->host_ag) {
    if (c->host_ag(c->host_ctx, h_send, h_recv, (int64_t)bytes) != 0) {
      pp_set_error("the application's all-gather callback failed");
      return PP_ERR_NCCL;
    }
    return PP_OK;
  }
  PP_NEED_NCCL(c);
  char *d_one, *d_all;
  PP_TRY(pp_dev_alloc(&d_one, bytes, s));
  PP_TRY(pp_dev_alloc(&d_all, bytes * (size_t)c->nranks, s));
  PP_CUDA(cudaMemcpyAsync(d_one, h_send, bytes, cudaMemcpyHostToDevice, s));
  PP_NCCL(g_nccl.AllGather(d_one, d_all, bytes, ncclUint8, c->comm, s));
  PP_CUDA(cudaMemcpyAsync(h_recv, d_all, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  pp_dev_free(d_one, s); pp_dev_free(d_all, s);
  return PP_OK;
}
pp_status boot_max(pp_comm* c, long long* v, cudaStream_t s) {
  std::vector<long long> all((size_t)c->nranks);
  PP_TRY(boot_allgather(c, v, all.data(), sizeof(long long), s));
  for (long long x : all) *v = x > *v ? x : *v;
  return PP_OK;
}
pp_status boot_min(pp_comm* c, int* v, cudaStream_t s) {       // also a barrier
  std::vector<int> all((size_t)c->nranks);
  PP_TRY(boot_allgather(c, v, all.data(), sizeof(int), s));
  for (int x : all) *v = x < *v ? x : *v;
  return PP_OK;
}

// device scratch of one call, released (stream-ordered) when the scope ends -- on error returns too
struct DevScope {
  cudaStream_t s;
  std::vector<void*> held;
  explicit DevScope(cudaStream_t s_) : s(s_) {}
  ~DevScope() { for (void* p : held) cudaFreeAsync(p, s); }
  template <class T>
  pp_status alloc(T** out, size_t n) {
    PP_TRY(pp_dev_alloc(out, n, s));
    held.push_back((void*)*out);
    return PP_OK;
  }
};
// ncclGroupStart / ncclGroupEnd as a scope: an early return between the two must still close the group
struct NcclGroup {
  bool open = false;
  pp_status begin() { PP_NCCL(g_nccl.GroupStart()); open = true; return PP_OK; }
  pp_status end() { open = false; PP_NCCL(g_nccl.GroupEnd()); return PP_OK; }
  ~NcclGroup() { if (open) g_nccl.GroupEnd(); }
};

pp_status init_nccl(pp_comm* c, const uint8_t id[128]) {
  const int nranks = c->nranks, rank = c->rank;
  PP_TRY(load_nccl());
  ncclUniqueId uid;
  memcpy(uid.internal, id, 128);
  PP_NCCL(g_nccl.CommInitRank(&c->comm, nranks, uid, rank));
  // NCCL connects point-to-point and collective channels lazily, on the first call that uses
  // them (tens of milliseconds per new peer): a migration that reaches a diagonal neighbour for
  // the first time in step 40 would stall that step.  Touch every peer and both collectives now.
  char* w = nullptr;
  PP_CUDA(cudaMalloc(&w, (size_t)(2 * nranks + 2) * 8));
  PP_CUDA(cudaMemset(w, 0, (size_t)(2 * nranks + 2) * 8));
  NcclGroup group1;
  PP_TRY(group1.begin());
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) continue;
    PP_NCCL(g_nccl.Send(w + 8 * p, 8, ncclUint8, p, c->comm, 0));
    PP_NCCL(g_nccl.Recv(w + 8 * (nranks + p), 8, ncclUint8, p, c->comm, 0));
  }
  PP_TRY(group1.end());
  PP_NCCL(g_nccl.AllReduce(w, w, 1, ncclInt64, ncclSum, c->comm, 0));
  PP_NCCL(g_nccl.AllGather(w + 8 * 2 * nranks, w, 2, ncclInt32, c->comm, 0));
  PP_CUDA(cudaStreamSynchronize(0));
  PP_CUDA(cudaFree(w));
  return PP_OK;
}
pp_comm* new_comm(int nranks, int rank) {
  pp_comm* c = new pp_comm();
  c->comm = nullptr; c->nranks = nranks; c->rank = rank;
  for (int p = 0; p < kMaxRanks; ++p) { c->p2p.peers.win[p] = nullptr; c->red.peers.win[p] = nullptr; }
  return c;
}
}  // namespace

extern "C" pp_status pp_comm_unique_id(uint8_t id_out[128]) {
  PP_REQUIRE(id_out, "null argument");
  PP_TRY(load_nccl());
  ncclUniqueId id;
  PP_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return PP_OK;
}

extern "C" pp_status pp_comm_create(int32_t nranks, int32_t rank, const uint8_t id[128], pp_comm** out) {
  PP_REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
  PP_REQUIRE(nranks == 1 || id, "a unique id is required for more than one rank");
  pp_comm* c = new_comm(nranks, rank);
  if (nranks > 1) {
    const pp_status st = init_nccl(c, id);
    if (st != PP_OK) { delete c; return st; }
  }
  *out = c;
  return PP_OK;
}

// The application brings its own bootstrap (the reference's world is MPI: support/ViewComm.h takes an
// MPI_Comm): `allgather` is called on the host, collectively, only while a communicator or one of its
// peer-memory windows is being set up.  use_nccl != 0: the NCCL id travels through it and the
// communicator is the same as pp_comm_create's.  use_nccl == 0: no NCCL at all -- migration and the
// comm-array reductions run over the peer-memory windows (CUDA IPC), which also works between several
// processes on ONE GPU (how the multi-rank tests run on a single-GPU box).
extern "C" pp_status pp_comm_create_hosted(int32_t nranks, int32_t rank, pp_host_allgather_fn allgather,
                                           void* ctx, int32_t use_nccl, pp_comm** out) {
  PP_REQUIRE(out && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
  PP_REQUIRE(nranks == 1 || allgather, "an all-gather callback is required for more than one rank");
  PP_REQUIRE(nranks <= kMaxRanks || use_nccl, "a communicator without NCCL holds at most 64 ranks");
  pp_comm* c = new_comm(nranks, rank);
  c->host_ag = allgather; c->host_ctx = ctx;
  if (nranks > 1 && use_nccl) {
    std::vector<uint8_t> mine(128, 0), all((size_t)nranks * 128);
    pp_status st = rank == 0 ? pp_comm_unique_id(mine.data()) : PP_OK;
    // the gather is entered by every rank, also when rank 0 failed (its id stays zero and NCCL rejects it)
    const pp_status st2 = boot_allgather(c, mine.data(), all.data(), 128, 0);
    if (st == PP_OK) st = st2;
    if (st == PP_OK) st = init_nccl(c, all.data());
    if (st != PP_OK) { delete c; return st; }
  }
  *out = c;
  return PP_OK;
}

extern "C" pp_status pp_comm_destroy(pp_comm* c) {
  if (!c) return PP_OK;
  if (c->p2p.local) {
    cudaDeviceSynchronize();
    for (int p = 0; p < c->nranks; ++p)
      if (p != c->rank && c->p2p.peers.win[p]) cudaIpcCloseMemHandle(c->p2p.peers.win[p]);
    cudaFree(c->p2p.local); cudaFree(c->p2p.dev); cudaFree(c->p2p.dev_stats);
    if (c->p2p.host_stats) cudaFreeHost(c->p2p.host_stats);
    if (c->p2p.ev) cudaEventDestroy(c->p2p.ev);
  }
  if (c->red.local) {
    cudaDeviceSynchronize();
    for (int p = 0; p < c->nranks; ++p)
      if (p != c->rank && c->red.peers.win[p]) cudaIpcCloseMemHandle(c->red.peers.win[p]);
    cudaFree(c->red.local); cudaFree(c->red.err);
  }
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  return PP_OK;
}
extern "C" int32_t pp_comm_size(const pp_comm* c) { return c ? c->nranks : -1; }
extern "C" int32_t pp_comm_rank(const pp_comm* c) { return c ? c->rank : -1; }

static pp_status nccl_type(int32_t dtype, ncclDataType_t* t, size_t* bytes) {
  switch (dtype) {
    case PP_INT32: *t = ncclInt32; *bytes = 4; return PP_OK;
    case PP_INT64: *t = ncclInt64; *bytes = 8; return PP_OK;
    case PP_FLOAT32: *t = ncclFloat32; *bytes = 4; return PP_OK;
    case PP_FLOAT64: *t = ncclFloat64; *bytes = 8; return PP_OK;
    default: pp_set_error("unknown pp_dtype %d", dtype); return PP_ERR_INVALID;
  }
}

namespace {
// the peer-memory reduction of pp_comm_array_reduce, in place on `arr`; *done = false: not available
pp_status red_try(pp_comm* c, void* arr, long n, int32_t dtype, int32_t op, cudaStream_t s, bool* done);
}
// PS_Comm_Allreduce (support/ViewComm_gpu.hpp:184-210) on device memory, in place allowed
extern "C" pp_status pp_comm_allreduce(pp_comm* c, const void* send, void* recv, int64_t count,
                                       int32_t dtype, int32_t op, pp_stream stream) {
  PP_REQUIRE(c && send && recv && count >= 0, "bad argument");
  ncclDataType_t t; size_t b;
  PP_TRY(nccl_type(dtype, &t, &b));
  if (c->nranks == 1) {
    if (send != recv) PP_CUDA(cudaMemcpyAsync(recv, send, b * count, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PP_OK;
  }
  PP_REQUIRE(op == PP_SUM || op == PP_MAX || op == PP_MIN, "unsupported reduction");
  if (!c->comm) {                    // hosted communicator without NCCL: the peer-memory reduction
    cudaStream_t s = (cudaStream_t)stream;
    if (send != recv) PP_CUDA(cudaMemcpyAsync(recv, send, b * count, cudaMemcpyDeviceToDevice, s));
    bool done = false;
    PP_TRY(red_try(c, recv, (long)count, dtype, op, s, &done));
    PP_REQUIRE(done || count == 0, "no transport: the peer-memory window could not be mapped and this "
                                   "communicator was created without NCCL");
    return PP_OK;
  }
  const ncclRedOp_t o = op == PP_SUM ? ncclSum : op == PP_MAX ? ncclMax : ncclMin;
  PP_NCCL(g_nccl.AllReduce(send, recv, (size_t)count, t, o, c->comm, (cudaStream_t)stream));
  return PP_OK;
}

// PS_Comm_Alltoall (support/ViewComm_gpu.hpp:130-160): `count` elements per peer
extern "C" pp_status pp_comm_alltoall(pp_comm* c, const void* send, void* recv, int64_t count,
                                      int32_t dtype, pp_stream stream) {
  PP_REQUIRE(c && send && recv && count >= 0, "bad argument");
  ncclDataType_t t; size_t b;
  PP_TRY(nccl_type(dtype, &t, &b));
  cudaStream_t s = (cudaStream_t)stream;
  if (c->nranks == 1) {
    if (send != recv) PP_CUDA(cudaMemcpyAsync(recv, send, b * count, cudaMemcpyDeviceToDevice, s));
    return PP_OK;
  }
  PP_NEED_NCCL(c);
  NcclGroup group2;
  PP_TRY(group2.begin());
  for (int p = 0; p < c->nranks; ++p) {
    PP_NCCL(g_nccl.Send((const char*)send + (size_t)p * count * b, (size_t)count, t, p, c->comm, s));
    PP_NCCL(g_nccl.Recv((char*)recv + (size_t)p * count * b, (size_t)count, t, p, c->comm, s));
  }
  PP_TRY(group2.end());
  return PP_OK;
}

// PS_Comm_Send / PS_Comm_Recv (support/ViewComm_gpu.hpp:13-60): blocking pair semantics are the
// caller's responsibility (NCCL matches a send with the peer's recv on the same communicator)
extern "C" pp_status pp_comm_send(pp_comm* c, const void* buf, int64_t count, int32_t dtype,
                                  int32_t peer, pp_stream stream) {
  PP_REQUIRE(c && c->nranks > 1 && (buf || count == 0) && peer >= 0 && peer < c->nranks, "bad argument");
  ncclDataType_t t; size_t b;
  PP_TRY(nccl_type(dtype, &t, &b));
  PP_NEED_NCCL(c);
  PP_NCCL(g_nccl.Send(buf, (size_t)count, t, peer, c->comm, (cudaStream_t)stream));
  return PP_OK;
}
extern "C" pp_status pp_comm_recv(pp_comm* c, void* buf, int64_t count, int32_t dtype,
                                  int32_t peer, pp_stream stream) {
  PP_REQUIRE(c && c->nranks > 1 && (buf || count == 0) && peer >= 0 && peer < c->nranks, "bad argument");
  ncclDataType_t t; size_t b;
  PP_TRY(nccl_type(dtype, &t, &b));
  PP_NEED_NCCL(c);
  PP_NCCL(g_nccl.Recv(buf, (size_t)count, t, peer, c->comm, (cudaStream_t)stream));
  return PP_OK;
}
extern "C" pp_status pp_comm_group_start(void) { PP_TRY(load_nccl()); PP_NCCL(g_nccl.GroupStart()); return PP_OK; }
extern "C" pp_status pp_comm_group_end(void) { PP_TRY(load_nccl()); PP_NCCL(g_nccl.GroupEnd()); return PP_OK; }

// ------------------------------------------------------------------------------------------
// Mesh::reduceCommArray for full-mesh PICparts (pumipic_comm.cpp:223-247) + BCAST_OP
// ------------------------------------------------------------------------------------------
namespace {
template <class T>
__global__ void k_mask_not_owned(T* a, const int* __restrict__ owner, int self, long nents, int nvals) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= nents * nvals) return;
  if (owner[i / nvals] != self) a[i] = T(0);
}
}  // namespace

namespace {
constexpr size_t kRedHdr = 1024;      // header bytes in front of A
__device__ __forceinline__ unsigned long long red_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// every block: wait until all ranks' flags show `epoch` (flags live in this rank's own window)
__device__ __forceinline__ void red_wait(const unsigned* flags, int nranks, unsigned epoch, int* err) {
  if (threadIdx.x < nranks) {
    const unsigned long long t0 = red_now_ns();
    while (*(volatile const unsigned*)(flags + threadIdx.x) != epoch) {
      if (red_now_ns() - t0 > 20000000000ull) { atomicExch(err, 1); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
}
__global__ void k_red_flag(P2PPeers peers, int self, int nranks, unsigned epoch, int which) {
  const int p = threadIdx.x;
  if (p >= nranks) return;
  __threadfence_system();
  P2PReduceHeader* h = reinterpret_cast<P2PReduceHeader*>(peers.win[p]);
  *(volatile unsigned*)(which == 0 ? &h->ready[self] : &h->done[self]) = epoch;
}
template <class T>
__device__ __forceinline__ T red_op(T a, T b, int op) {
  return op == PP_SUM ? a + b : op == PP_MAX ? (a < b ? b : a) : (b < a ? b : a);
}
// slice `self` of all ranks' inputs, reduced in ascending rank order, stored into all ranks' results
template <class T>
__global__ void __launch_bounds__(256) k_red_slice(P2PPeers peers, int self, int nranks, long n, size_t cap_bytes,
                                                    unsigned epoch, int op, int* err) {
  const P2PReduceHeader* h = reinterpret_cast<const P2PReduceHeader*>(peers.win[self]);
  red_wait(h->ready, nranks, epoch, err);
  const long per = ((n + nranks - 1) / nranks + 1) & ~1l;       // even: slices stay 16-byte aligned for doubles
  const long lo = min(n, per * self), hi = min(n, lo + per);
  if (sizeof(T) == 8 && ((hi - lo) & 1) == 0) {
    // two elements per 16-byte access (NVLink packets of 16 bytes and more)
    typedef typename std::conditional<std::is_same<T, double>::value, double2, longlong2>::type T2;
    const long lo2 = lo >> 1, hi2 = hi >> 1;
    for (long i = lo2 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < hi2; i += (long)gridDim.x * blockDim.x) {
      T2 acc = reinterpret_cast<const T2*>(peers.win[0] + kRedHdr)[i];
      for (int p = 1; p < nranks; ++p) {
        const T2 v = reinterpret_cast<const T2*>(peers.win[p] + kRedHdr)[i];
        acc.x = red_op(acc.x, v.x, op); acc.y = red_op(acc.y, v.y, op);
      }
      for (int q = 0; q < nranks; ++q) reinterpret_cast<T2*>(peers.win[q] + kRedHdr + cap_bytes)[i] = acc;
    }
    return;
  }
  for (long i = lo + blockIdx.x * (long)blockDim.x + threadIdx.x; i < hi; i += (long)gridDim.x * blockDim.x) {
    T acc = reinterpret_cast<const T*>(peers.win[0] + kRedHdr)[i];
    for (int p = 1; p < nranks; ++p) acc = red_op(acc, reinterpret_cast<const T*>(peers.win[p] + kRedHdr)[i], op);
    for (int q = 0; q < nranks; ++q) reinterpret_cast<T*>(peers.win[q] + kRedHdr + cap_bytes)[i] = acc;
  }
}
template <class T>
__global__ void __launch_bounds__(256) k_red_copy_out(T* __restrict__ arr, const char* local, long n, size_t cap_bytes,
                                                       int nranks, unsigned epoch, int* err) {
  const P2PReduceHeader* h = reinterpret_cast<const P2PReduceHeader*>(local);
  red_wait(h->done, nranks, epoch, err);
  const T* B = reinterpret_cast<const T*>(local + kRedHdr + cap_bytes);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) arr[i] = B[i];
}

// collective: (re)create the reduction window for arrays of want_bytes (every rank calls it with the
// same size: the comm arrays of an all-reduce have the same length everywhere)
pp_status red_setup(pp_comm* c, size_t want_bytes, cudaStream_t s) {
  P2PReduce& w = c->red;
  const int R = c->nranks, me = c->rank;
  if (w.local) {                       // grow: nobody touches the old windows once every device is idle
    PP_CUDA(cudaDeviceSynchronize());
    for (int p = 0; p < R; ++p)
      if (p != me && w.peers.win[p]) { cudaIpcCloseMemHandle(w.peers.win[p]); w.peers.win[p] = nullptr; }
    int closed = 1;                    // barrier: all mappings are closed before an owner frees its window
    PP_TRY(boot_min(c, &closed, s));
    cudaFree(w.local); w.local = nullptr;
    if (w.err) { cudaFree(w.err); w.err = nullptr; }
    w.ok = false;
    want_bytes += want_bytes / 2;      // head room: do not come back for every few per cent
  }
  w.tried = true;
  const char* env = getenv("PUMIPIC_P2P");
  int ok = (R <= kMaxRanks && g_p2p_enable && !(env && env[0] == '0')) ? 1 : 0;
  long long h_sz = (long long)((want_bytes + 255) & ~(size_t)255);
  PP_TRY(boot_max(c, &h_sz, s));
  w.cap_bytes = (size_t)h_sz;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc((void**)&w.local, kRedHdr + 2 * w.cap_bytes) != cudaSuccess) { ok = 0; w.local = nullptr; cudaGetLastError(); }
  if (ok) {
    PP_CUDA(cudaMemset(w.local, 0, kRedHdr));
    PP_CUDA(cudaDeviceSynchronize());   // the header is zero before any peer can learn the handle
    if (cudaIpcGetMemHandle(&mine, w.local) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  Msg m; m.h = mine; m.ok = ok; m.pad[0] = m.pad[1] = m.pad[2] = 0;
  std::vector<Msg> all((size_t)R);
  PP_TRY(boot_allgather(c, &m, all.data(), sizeof(Msg), s));
  for (int p = 0; p < R; ++p) ok &= all[(size_t)p].ok;
  if (ok) {
    w.peers.win[me] = w.local;
    for (int p = 0; p < R && ok; ++p) {
      if (p == me) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[(size_t)p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
      w.peers.win[p] = (char*)q;
    }
  }
  PP_TRY(boot_min(c, &ok, s));         // everybody mapped everybody (also a barrier: all headers are zero)
  if (ok) {
    PP_CUDA(cudaMalloc((void**)&w.err, sizeof(int)));
    PP_CUDA(cudaMemset(w.err, 0, sizeof(int)));
  } else if (w.local) {
    for (int p = 0; p < R; ++p)
      if (p != me && w.peers.win[p]) { cudaIpcCloseMemHandle(w.peers.win[p]); w.peers.win[p] = nullptr; }
    cudaFree(w.local);
    w.local = nullptr;
  }
  w.ok = ok != 0;
  return PP_OK;
}

template <class T>
pp_status red_run(pp_comm* c, T* arr, long n, int op, cudaStream_t s) {
  P2PReduce& w = c->red;
  const int R = c->nranks, me = c->rank;
  const unsigned epoch = ++w.epoch;
  PP_CUDA(cudaMemcpyAsync(w.local + kRedHdr, arr, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  k_red_flag<<<1, kMaxRanks, 0, s>>>(w.peers, me, R, epoch, 0);
  const long per = (n + R - 1) / R;
  const int grid = (int)std::max<long>(1, std::min<long>((per + 255) / 256, 148 * 8));
  k_red_slice<T><<<grid, 256, 0, s>>>(w.peers, me, R, n, w.cap_bytes, epoch, op, w.err);
  k_red_flag<<<1, kMaxRanks, 0, s>>>(w.peers, me, R, epoch, 1);
  const int grid2 = (int)std::max<long>(1, std::min<long>((n + 255) / 256, 148 * 8));
  k_red_copy_out<T><<<grid2, 256, 0, s>>>(arr, w.local, n, w.cap_bytes, R, epoch, w.err);
  PP_KERNEL_CHECK();
  return PP_OK;
}
}  // namespace

namespace {
pp_status red_try(pp_comm* c, void* arr, long n, int32_t dtype, int32_t op, cudaStream_t s, bool* done) {
  *done = false;
  const size_t esz = (dtype == PP_INT32 || dtype == PP_FLOAT32) ? 4 : 8;
  if (n <= 0 || !(op == PP_SUM || op == PP_MAX || op == PP_MIN)) return PP_OK;
  if (!c->red.tried || (c->red.ok && (size_t)n * esz > c->red.cap_bytes)) PP_TRY(red_setup(c, (size_t)n * esz, s));
  if (!c->red.ok || (size_t)n * esz > c->red.cap_bytes) return PP_OK;
  *done = true;
  switch (dtype) {
    case PP_INT32: return red_run(c, (int*)arr, n, op, s);
    case PP_INT64: return red_run(c, (long long*)arr, n, op, s);
    case PP_FLOAT32: return red_run(c, (float*)arr, n, op, s);
    case PP_FLOAT64: return red_run(c, (double*)arr, n, op, s);
    default: *done = false; return PP_OK;
  }
}
}  // namespace

extern "C" pp_status pp_comm_array_reduce(pp_comm* c, void* comm_array, int64_t nents, int32_t nvals,
                                          int32_t dtype, int32_t op, const int32_t* ent_owner,
                                          pp_stream stream) {
  PP_REQUIRE(c && comm_array && nents >= 0 && nvals >= 1, "bad argument");
  if (c->nranks == 1) return PP_OK;          // pumipic_comm.cpp:232-233
  cudaStream_t s = (cudaStream_t)stream;
  int32_t nccl_op = op;
  if (op == PP_BCAST) {
    // owner's value wins: zero every copy that is not the owner's, then sum (adds exact zeros)
    PP_REQUIRE(ent_owner, "BCAST needs the entity owners");
    const long n = nents * nvals;
    const int g = pp_div_up(n, kBlock);
    if (n > 0) {
      if (dtype == PP_FLOAT64) k_mask_not_owned<<<g, kBlock, 0, s>>>((double*)comm_array, ent_owner, c->rank, nents, nvals);
      else if (dtype == PP_FLOAT32) k_mask_not_owned<<<g, kBlock, 0, s>>>((float*)comm_array, ent_owner, c->rank, nents, nvals);
      else if (dtype == PP_INT32) k_mask_not_owned<<<g, kBlock, 0, s>>>((int*)comm_array, ent_owner, c->rank, nents, nvals);
      else k_mask_not_owned<<<g, kBlock, 0, s>>>((long long*)comm_array, ent_owner, c->rank, nents, nvals);
      PP_KERNEL_CHECK();
    }
    nccl_op = PP_SUM;
  }
  // peer-memory reduction (default between GPUs with peer access): reduce-scatter + all-gather by
  // direct NVLink loads / stores, fixed rank order; the window is sized by the first call
  bool done = false;
  PP_TRY(red_try(c, comm_array, (long)(nents * nvals), dtype, nccl_op, s, &done));
  if (done) return PP_OK;
  return pp_comm_allreduce(c, comm_array, comm_array, nents * nvals, dtype, nccl_op, stream);
}

// ------------------------------------------------------------------------------------------
// migrate (SCS_migrate.h:5-221, identical algorithm in CSR_migrate.hpp / dps / cabm)
// ------------------------------------------------------------------------------------------
namespace {
__global__ void k_count_dest(PsView v, const int* __restrict__ new_proc, const int* __restrict__ new_elem,
                             int self, int nranks, int* cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int dest = -1;
  if (s < v.capacity && ((__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u)) {
    const int p = new_proc[s];
    if (p != self && p >= 0 && p < nranks && new_elem[s] >= 0) dest = p;   // deleted particles are not sent
  }
  const unsigned grp = __match_any_sync(0xffffffffu, dest);
  if (dest >= 0 && (threadIdx.x & 31) == (__ffs(grp) - 1)) atomicAdd(cnt + dest, __popc(grp));
}

struct PackTable {
  int n;
  const char* src[16];
  int bytes[16];
  int ncomp[16];
};

// wire format of one peer block holding n particles (all sections 8-byte aligned):
//   int64 gid[n] | member0: [ncomp0][n] scalars | member1 ... (LayoutLeft per member, as the
//   reference's per-member messages, MemberTypeLibraries.h:272-279)
__host__ __device__ inline size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }
__host__ __device__ inline size_t block_bytes(const PackTable& t, size_t n) {
  size_t b = 8 * n;
  for (int k = 0; k < t.n; ++k) b += align8((size_t)t.bytes[k] * t.ncomp[k] * n);
  return b;
}

__global__ void k_pack(PsView v, const int* __restrict__ new_proc, int* new_elem,
                       const long long* __restrict__ elem_gids, int self, int nranks,
                       const int* __restrict__ send_cnt, const size_t* __restrict__ peer_byte_off,
                       int* cursor, PackTable t, long stride, char* sendbuf) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity) return;
  if (!((__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u)) return;
  const int p = new_proc[s];
  if (p == self || p < 0 || p >= nranks) return;
  const int e = new_elem[s];
  if (e < 0) return;                       // deleted by the caller: dropped here, not sent
  const int i = atomicAdd(cursor + p, 1);
  const size_t n = (size_t)send_cnt[p];
  char* blk = sendbuf + peer_byte_off[p];
  ((long long*)blk)[i] = e < 0 ? -1 : (elem_gids ? elem_gids[e] : (long long)e);
  size_t off = 8 * n;
  for (int k = 0; k < t.n; ++k) {
    const int sb = t.bytes[k];
    for (int c = 0; c < t.ncomp[k]; ++c) {
      const char* a = t.src[k] + ((size_t)c * stride + s) * sb;
      char* b = blk + off + ((size_t)c * n + i) * sb;
      if (sb == 8) *(double*)b = *(const double*)a;
      else if (sb == 4) *(int*)b = *(const int*)a;
      else for (int q = 0; q < sb; ++q) b[q] = a[q];
    }
    off += align8((size_t)sb * t.ncomp[k] * n);
  }
  new_elem[s] = -1;   // removeSentParticles (SCS_migrate.h:190-196)
}

// received peer blocks -> [ncomp][n_total] member arrays + local element ids
__global__ void k_unpack(const char* __restrict__ recvbuf, const size_t* __restrict__ peer_byte_off,
                         const int* __restrict__ recv_cnt, const int* __restrict__ recv_off, int nranks,
                         int n_total, int row_len, PackTable t, char* const* dst,
                         const long long* __restrict__ sorted_gid,
                         const int* __restrict__ sorted_lid, int ne, int* elems_out, int* bad) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_total) return;
  int p = 0;
  while (p + 1 < nranks && recv_off[p + 1] <= j) ++p;
  const int i = j - recv_off[p];
  const size_t n = (size_t)recv_cnt[p];
  const char* blk = recvbuf + peer_byte_off[p];
  const long long gid = ((const long long*)blk)[i];
  int lid = -1;
  if (sorted_gid) {                    // gid -> lid (replaces Kokkos::UnorderedMap, SCS_migrate.h:181-187)
    int lo = 0, hi = ne - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const long long g = sorted_gid[mid];
      if (g == gid) { lid = sorted_lid[mid]; break; }
      if (g < gid) lo = mid + 1; else hi = mid - 1;
    }
  } else if (gid >= 0 && gid < ne) {
    lid = (int)gid;
  }
  if (lid < 0) *bad = 1;
  elems_out[j] = lid;
  size_t off = 8 * n;
  for (int k = 0; k < t.n; ++k) {
    const int sb = t.bytes[k];
    for (int c = 0; c < t.ncomp[k]; ++c) {
      const char* a = blk + off + ((size_t)c * n + i) * sb;
      char* b = dst[k] + ((size_t)c * row_len + j) * sb;
      if (sb == 8) *(double*)b = *(const double*)a;
      else if (sb == 4) *(int*)b = *(const int*)a;
      else for (int q = 0; q < sb; ++q) b[q] = a[q];
    }
    off += align8((size_t)sb * t.ncomp[k] * n);
  }
}

__global__ void k_iota(int* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
}  // namespace

extern "C" pp_status pp_ps_rebuild(pp_ps* ps, const int32_t* new_element, int32_t n_new,
                                   const int32_t* new_particle_elements,
                                   const void* const* new_particle_info, pp_stream stream_);

// ------------------------------------------------------------------------------------------
// migration over the peer-memory window
// ------------------------------------------------------------------------------------------
namespace {
// layout of the window's device scratch (ints)
struct P2PScratch {
  int* cursor; int* overflow; int* err; int* n_in; int* n_total; int* recv_off; int* recv_cnt;
};
P2PScratch p2p_scratch(int* base, int R) {
  P2PScratch x;
  x.cursor = base; x.overflow = base + R; x.err = x.overflow + 1; x.n_in = x.err + 1; x.n_total = x.n_in + 1;
  x.recv_off = x.n_total + 1; x.recv_cnt = x.recv_off + R + 1;
  return x;
}
__host__ __device__ inline size_t p2p_rec_bytes(const PackTable& t) {   // multiple of 16: moved in 16-byte pieces
  size_t b = 8;
  for (int k = 0; k < t.n; ++k) b += align8((size_t)t.bytes[k] * t.ncomp[k]);
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t p2p_seg_offset(int parity, int sender, int R, size_t seg_bytes) {
  return sizeof(P2PHeader) + ((size_t)parity * R + sender) * seg_bytes;
}

// record of one particle: int64 gid | member 0 components | member 1 ... (members 8-byte aligned)
// A block owns kPackSlots consecutive slots: it counts its leaving particles per destination in
// shared memory, reserves their places in the destinations' segments with ONE global atomic per
// destination (thousands of lanes adding to the same few cursors would serialise in L2), then
// stores the records over NVLink.
constexpr int kPackThreads = 256, kPackPerThread = 8, kPackSlots = kPackThreads * kPackPerThread;
constexpr int kPackCoopMin = 32;    // leavers per block (of 2048 slots) from which the warp-cooperative stores pay
__global__ void __launch_bounds__(kPackThreads) k_p2p_pack(
    PsView v, const int* __restrict__ new_proc, int* new_elem, const long long* __restrict__ elem_gids, int self,
    int nranks, int* cursor, int* overflow, P2PPeers peers, int parity, size_t seg_bytes, int seg_cap,
    int rec_bytes, PackTable t, long stride, int debug) {
  __shared__ int s_cnt[kMaxRanks], s_base[kMaxRanks];
  __shared__ int s_total;
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  // slots of thread t: 4 consecutive ones per 16-byte load of new_proc, kPackPerThread / 4 loads
  const long sb0 = (long)blockIdx.x * kPackSlots + 4 * threadIdx.x;
  const bool vec = ((size_t)new_proc & 15) == 0;
  int dest[kPackPerThread], pos[kPackPerThread];
#pragma unroll
  for (int g = 0; g < kPackPerThread / 4; ++g) {
    const long sg = sb0 + (long)g * 4 * kPackThreads;
    int pv[4] = {self, self, self, self};
    if (vec && sg + 3 < v.capacity) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(new_proc + sg));
      pv[0] = q.x; pv[1] = q.y; pv[2] = q.z; pv[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (sg + j < v.capacity) pv[j] = new_proc[sg + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * g + j;
      const long s = sg + j;
      const int p = pv[j];
      dest[k] = -1;
      // most slots stay (p == self): the mask and the element are only read for the others; a
      // particle the caller deleted (element -1) is dropped here, not sent
      if (p != self && p >= 0 && p < nranks && s < v.capacity &&
          ((__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u) && new_elem[s] >= 0)
        dest[k] = p;
      if (debug == 2) dest[k] = -1;
      if (dest[k] >= 0) pos[k] = atomicAdd(&s_cnt[dest[k]], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < nranks) {
    const int c = s_cnt[threadIdx.x];
    s_base[threadIdx.x] = c ? atomicAdd(cursor + threadIdx.x, c) : 0;
    if (c) atomicAdd(&s_total, c);
  }
  __syncthreads();
  if (s_total == 0) return;
  // Bulk migrations (a block with many leavers): records leave the GPU as whole records -- the WARP
  // stores one leaving particle at a time, lane i carrying scalar i of the record (gid, then the
  // members' components), so the record's bytes cross NVLink as one contiguous run instead of one
  // 8-byte store per scalar from a single lane (ps_combo160's migrate series, 5.5 M records of 168
  // bytes per step: 32.8 -> 17.4 ms).  A PIC step's handful of leavers per block (0.3 % of the slots)
  // is cheaper the plain way below: the warp-wide loop costs every thread its ballots.
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  int nsc = 1;
  for (int m = 0; m < t.n; ++m) nsc += t.ncomp[m];
  if (nsc <= 2 * 32 && s_total >= kPackCoopMin && debug != 3) {
    // scalar `i` of a record: source column base, byte offset in the record, scalar size (0 = none, -1 = gid)
    auto describe = [&](int i, const char*& base, int& off, int& sb) {
      base = nullptr; off = 0; sb = 0;
      if (i == 0) { sb = -1; return; }
      int idx = 1, o = 8;
      for (int m = 0; m < t.n; ++m) {
        if (i < idx + t.ncomp[m]) {
          const int c = i - idx;
          sb = t.bytes[m];
          base = t.src[m] + (size_t)c * stride * sb;
          off = o + c * sb;
          return;
        }
        idx += t.ncomp[m];
        o += (int)align8((size_t)t.bytes[m] * t.ncomp[m]);
      }
    };
    const char *b0, *b1;
    int o0, o1, z0, z1;
    describe(lane, b0, o0, z0);
    describe(lane + 32, b1, o1, z1);
    auto put = [&](const char* base, int off, int sb, char* rec, long sl, long long gid) {
      if (sb == -1) *(long long*)rec = gid;
      else if (sb == 8) *(double*)(rec + off) = *(const double*)(base + sl * 8);
      else if (sb == 4) *(int*)(rec + off) = *(const int*)(base + sl * 4);
      else for (int q = 0; q < sb; ++q) rec[off + q] = base[sl * sb + q];
    };
#pragma unroll
    for (int k = 0; k < kPackPerThread; ++k) {
      const long s = sb0 + (long)(k >> 2) * 4 * kPackThreads + (k & 3);
      const int p = dest[k];
      bool go = p >= 0;
      int i = 0;
      if (go) {
        i = s_base[p] + pos[k];
        if (i >= seg_cap) { atomicAdd(overflow, 1); go = false; }   // no room: stays on this rank this step
      }
      long long gid = 0;
      char* rec = nullptr;
      if (go) {
        const int e = new_elem[s];
        gid = elem_gids ? elem_gids[e] : (long long)e;
        rec = peers.win[debug == 1 ? self : p] + p2p_seg_offset(parity, self, nranks, seg_bytes) + (size_t)i * rec_bytes;
      }
      unsigned todo = __ballot_sync(full, go);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const long sl = __shfl_sync(full, s, src);
        const long long gl = __shfl_sync(full, gid, src);
        char* rl = reinterpret_cast<char*>(__shfl_sync(full, reinterpret_cast<unsigned long long>(rec), src));
        if (z0) put(b0, o0, z0, rl, sl, gl);
        if (z1) put(b1, o1, z1, rl, sl, gl);
      }
      if (go) new_elem[s] = -1;                      // removeSentParticles (SCS_migrate.h:190-196)
    }
    return;
  }
  // few leavers in this block, or more than 64 scalars per record: one lane stores its own record
#pragma unroll
  for (int k = 0; k < kPackPerThread; ++k) {
    if (dest[k] < 0) continue;
    const long s = sb0 + (long)(k >> 2) * 4 * kPackThreads + (k & 3);
    const int p = dest[k];
    const int i = s_base[p] + pos[k];
    if (i >= seg_cap) { atomicAdd(overflow, 1); continue; }   // no room: stays on this rank this step
    const int e = new_elem[s];
    char* rec = peers.win[debug == 1 ? self : p] + p2p_seg_offset(parity, self, nranks, seg_bytes) + (size_t)i * rec_bytes;
    *(long long*)rec = elem_gids ? elem_gids[e] : (long long)e;
    size_t off = 8;
    for (int m = 0; m < t.n; ++m) {
      const int sb = t.bytes[m];
      for (int c = 0; c < t.ncomp[m]; ++c) {
        const char* a = t.src[m] + ((size_t)c * stride + s) * sb;
        char* b = rec + off + (size_t)c * sb;
        if (sb == 8) *(double*)b = *(const double*)a;
        else if (sb == 4) *(int*)b = *(const int*)a;
        else for (int q = 0; q < sb; ++q) b[q] = a[q];
      }
      off += align8((size_t)sb * t.ncomp[m]);
    }
    new_elem[s] = -1;                      // removeSentParticles (SCS_migrate.h:190-196)
  }
}
// after the pack kernel has completed (its stores have landed): counts, then the step flag
__global__ void k_p2p_publish(const int* __restrict__ cursor, const int* __restrict__ overflow, int nranks,
                              int self, P2PPeers peers, int parity, unsigned epoch, int seg_cap,
                              long long* stats) {
  const int p = threadIdx.x;
  __shared__ long long sent;
  if (p == 0) sent = 0;
  __syncthreads();
  if (p < nranks && p != self) {
    const int cnt = min(cursor[p], seg_cap);
    atomicAdd((unsigned long long*)&sent, (unsigned long long)cnt);
    P2PHeader* h = reinterpret_cast<P2PHeader*>(peers.win[p]);
    *(volatile int*)&h->counts[parity][self] = cnt;
    __threadfence_system();
    *(volatile unsigned*)&h->flags[parity][self] = epoch;
  }
  __syncthreads();
  if (p == 0) { stats[0] = sent; stats[2] = *overflow; }
}
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin on this rank's header until every sender has published step `epoch`
__global__ void k_p2p_wait(P2PHeader* local_hdr, int nranks, int self, int parity, unsigned epoch,
                           int* recv_cnt, int* recv_off, int* n_in, int* n_total, int n_user, int* err,
                           long long* stats) {
  __shared__ int cnt[kMaxRanks];
  const int p = threadIdx.x;
  if (p < nranks) {
    int c = 0;
    if (p != self) {
      const unsigned long long t0 = p2p_now_ns();
      bool ok = true;
      while (*(volatile unsigned*)&local_hdr->flags[parity][p] != epoch) {
        if (p2p_now_ns() - t0 > 20000000000ull) { ok = false; break; }     // 20 s: a peer is gone
        __nanosleep(200);
      }
      __threadfence_system();
      if (ok) c = *(volatile int*)&local_hdr->counts[parity][p];
      else atomicExch(err, 1);
    }
    cnt[p] = c;
  }
  __syncthreads();
  if (p == 0) {
    int tot = 0;
    for (int r = 0; r < nranks; ++r) { recv_cnt[r] = cnt[r]; recv_off[r] = tot; tot += cnt[r]; }
    recv_off[nranks] = tot;
    *n_in = tot;
    *n_total = tot + n_user;
    stats[1] = tot;
    stats[3] = *err;
  }
}
// received records -> [ncomp][ld] member arrays + local element ids (grid-stride, count on device)
__global__ void k_p2p_unpack(const char* __restrict__ win, int parity, int nranks, size_t seg_bytes,
                             int rec_bytes, const int* __restrict__ recv_off, const int* __restrict__ n_in,
                             long ld, PackTable t, char* const* dst, const long long* __restrict__ sorted_gid,
                             const int* __restrict__ sorted_lid, int ne, int* elems_out) {
  const int n = *n_in;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    int p = 0;
    while (p + 1 < nranks && recv_off[p + 1] <= j) ++p;
    const char* rec = win + p2p_seg_offset(parity, p, nranks, seg_bytes) + (size_t)(j - recv_off[p]) * rec_bytes;
    const long long gid = *(const long long*)rec;
    int lid = -1;
    if (sorted_gid) {                    // gid -> lid (replaces Kokkos::UnorderedMap, SCS_migrate.h:181-187)
      int lo = 0, hi = ne - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const long long g = sorted_gid[mid];
        if (g == gid) { lid = sorted_lid[mid]; break; }
        if (g < gid) lo = mid + 1; else hi = mid - 1;
      }
    } else if (gid >= 0 && gid < ne) {
      lid = (int)gid;
    }
    elems_out[j] = lid;                  // -1 (unknown gid) makes the rebuild fail with an error
    size_t off = 8;
    for (int k = 0; k < t.n; ++k) {
      const int sb = t.bytes[k];
      for (int c = 0; c < t.ncomp[k]; ++c) {
        const char* a = rec + off + (size_t)c * sb;
        char* b = dst[k] + ((size_t)c * ld + j) * sb;
        if (sb == 8) *(double*)b = *(const double*)a;
        else if (sb == 4) *(int*)b = *(const int*)a;
        else for (int q = 0; q < sb; ++q) b[q] = a[q];
      }
      off += align8((size_t)sb * t.ncomp[k]);
    }
  }
}
// the caller's own new particles follow the received ones
__global__ void k_p2p_append(const int* __restrict__ n_in, int n_user, const int* __restrict__ user_elems,
                             PackTable user, long ld, char* const* dst, int* elems_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_user) return;
  const long at = (long)*n_in + j;
  elems_out[at] = user_elems[j];
  for (int k = 0; k < user.n; ++k) {
    const int sb = user.bytes[k];
    for (int c = 0; c < user.ncomp[k]; ++c) {
      const char* a = user.src[k] + ((size_t)c * n_user + j) * sb;
      char* b = dst[k] + ((size_t)c * ld + at) * sb;
      for (int q = 0; q < sb; ++q) b[q] = a[q];
    }
  }
}

// one-time, collective: allocate this rank's window, exchange IPC handles, map the peers
pp_status p2p_setup(pp_comm* c, size_t want_seg_bytes, cudaStream_t s) {
  P2PWindow& w = c->p2p;
  w.tried = true;
  const int R = c->nranks, me = c->rank;
  const char* env = getenv("PUMIPIC_P2P");            // PUMIPIC_P2P=0: NCCL path (A/B, tests)
  int ok = (R <= kMaxRanks && g_p2p_enable && !(env && env[0] == '0')) ? 1 : 0;
  // one segment size for every rank: the largest any rank asks for (senders address the receivers'
  // windows with it)
  if (!w.seg_set && want_seg_bytes > w.seg_bytes) w.seg_bytes = (want_seg_bytes + 15) & ~(size_t)15;
  {
    long long h_sz = (long long)w.seg_bytes;
    PP_TRY(boot_max(c, &h_sz, s));
    w.seg_bytes = (size_t)h_sz;
  }
  const size_t bytes = sizeof(P2PHeader) + 2 * (size_t)R * w.seg_bytes;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    if (cudaMalloc((void**)&w.local, bytes) != cudaSuccess) { ok = 0; w.local = nullptr; cudaGetLastError(); }
  }
  if (ok) {
    PP_CUDA(cudaMemset(w.local, 0, sizeof(P2PHeader)));
    PP_CUDA(cudaDeviceSynchronize());   // the header is zero before any peer can learn the handle
    if (cudaIpcGetMemHandle(&mine, w.local) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  // handles (and whether this rank got that far) to everybody
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Msg) % 8 == 0, "message size");
  Msg m; m.h = mine; m.ok = ok; m.pad[0] = m.pad[1] = m.pad[2] = 0;
  std::vector<Msg> all((size_t)R);
  PP_TRY(boot_allgather(c, &m, all.data(), sizeof(Msg), s));
  for (int p = 0; p < R; ++p) ok &= all[(size_t)p].ok;
  if (ok) {
    w.peers.win[me] = w.local;
    for (int p = 0; p < R && ok; ++p) {
      if (p == me) continue;
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[(size_t)p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        ok = 0; cudaGetLastError();
      }
      w.peers.win[p] = (char*)q;
    }
  }
  // everybody must have mapped everybody (also a barrier: no rank writes before all headers are zero)
  PP_TRY(boot_min(c, &ok, s));
  if (ok) {
    PP_CUDA(cudaMalloc((void**)&w.dev, sizeof(int) * (3 * (size_t)R + 8)));
    PP_CUDA(cudaMalloc((void**)&w.dev_stats, 4 * sizeof(long long)));
    PP_CUDA(cudaHostAlloc((void**)&w.host_stats, 4 * sizeof(long long), cudaHostAllocDefault));
    PP_CUDA(cudaEventCreateWithFlags(&w.ev, cudaEventDisableTiming));
  } else if (w.local) {
    for (int p = 0; p < R; ++p)
      if (p != me && w.peers.win[p]) { cudaIpcCloseMemHandle(w.peers.win[p]); w.peers.win[p] = nullptr; }
    cudaFree(w.local);
    w.local = nullptr;
  }
  w.ok = ok != 0;
  return PP_OK;
}

pp_status build_sorted_gids(pp_ps* ps, cudaStream_t s) {
  if (!ps->elem_gids || ps->sorted_gid) return PP_OK;
  // lazily build the sorted gid table (createGlobalMapping, SCS_buildFns.h:101-112)
  long long* keys_in = (long long*)ps->elem_gids;
  int* vals_in;
  PP_TRY(pp_dev_alloc(&vals_in, ps->nelems, s));
  PP_TRY(pp_dev_alloc(&ps->sorted_gid, ps->nelems, s));
  PP_TRY(pp_dev_alloc(&ps->sorted_lid, ps->nelems, s));
  k_iota<<<pp_div_up(ps->nelems, kBlock), kBlock, 0, s>>>(vals_in, ps->nelems);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, (long long*)ps->sorted_gid, vals_in, ps->sorted_lid,
                                  ps->nelems, 0, 64, s);
  char* tmp;
  PP_TRY(pp_dev_alloc(&tmp, tb, s));
  PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, (long long*)ps->sorted_gid, vals_in, ps->sorted_lid,
                                          ps->nelems, 0, 64, s));
  pp_dev_free(tmp, s); pp_dev_free(vals_in, s);
  return PP_OK;
}

pp_status migrate_p2p(pp_ps* ps, pp_comm* comm, int32_t* new_element, const int32_t* new_process, int32_t n_user,
                      const int32_t* user_elems, const void* const* user_info, pp_migrate_stats* stats_host,
                      const PackTable& pt, cudaStream_t s) {
  P2PWindow& w = comm->p2p;
  const int R = comm->nranks, me = comm->rank;
  const int rec_bytes = (int)p2p_rec_bytes(pt);
  const int seg_cap = (int)std::min<size_t>(w.seg_bytes / rec_bytes, 0x3fffffff / R);
  const unsigned epoch = ++w.epoch;
  const int parity = (int)(epoch & 1u);
  P2PScratch x = p2p_scratch(w.dev, R);
  PP_CUDA(cudaMemsetAsync(w.dev, 0, sizeof(int) * (R + 2), s));       // cursor, overflow, err
  PPTimeScope* t_x = new PPTimeScope(s, "migration pack + peer stores");
  if (ps->capacity > 0) {
    PP_TIME(s, "migration pack kernel");
    static const int dbg = getenv("PUMIPIC_PACK_DEBUG") ? atoi(getenv("PUMIPIC_PACK_DEBUG")) : 0;   // timing experiments only
    k_p2p_pack<<<pp_div_up(ps->capacity, kPackSlots), kPackThreads, 0, s>>>(
        ps->view(), new_process, new_element, (const long long*)ps->elem_gids, me, R, x.cursor, x.overflow,
        w.peers, parity, w.seg_bytes, seg_cap, rec_bytes, pt, ps->stride, dbg);
  }
  k_p2p_publish<<<1, kMaxRanks, 0, s>>>(x.cursor, x.overflow, R, me, w.peers, parity, epoch, seg_cap, w.dev_stats);
  delete t_x;
  t_x = new PPTimeScope(s, "migration wait for peers");
  k_p2p_wait<<<1, kMaxRanks, 0, s>>>((P2PHeader*)w.local, R, me, parity, epoch, x.recv_cnt, x.recv_off, x.n_in,
                                     x.n_total, n_user, x.err, w.dev_stats);
  delete t_x;
  PP_KERNEL_CHECK();
  PP_TIME(s, "migration unpack + rebuild");
  // new-particle arrays of the rebuild: rows of `ld` slots, received particles first
  const long ld = (long)(R - 1) * seg_cap + n_user;
  std::vector<char*> in_data(ps->nmembers, nullptr);
  int* in_elems;
  char** d_dst;
  PP_TRY(pp_dev_alloc(&in_elems, (size_t)ld + 1, s));
  PP_TRY(pp_dev_alloc(&d_dst, ps->nmembers, s));
  for (int k = 0; k < ps->nmembers; ++k)
    PP_TRY(pp_dev_alloc(&in_data[k], (size_t)pt.bytes[k] * pt.ncomp[k] * (size_t)ld + 8, s));
  PP_CUDA(cudaMemcpyAsync(d_dst, in_data.data(), sizeof(char*) * ps->nmembers, cudaMemcpyHostToDevice, s));
  PP_TRY(build_sorted_gids(ps, s));
  k_p2p_unpack<<<256, kBlock, 0, s>>>(w.local, parity, R, w.seg_bytes, rec_bytes, x.recv_off, x.n_in, ld, pt, d_dst,
                                      (const long long*)ps->sorted_gid, ps->sorted_lid, ps->nelems, in_elems);
  if (n_user > 0) {
    PackTable ut = pt;
    for (int k = 0; k < ut.n; ++k) ut.src[k] = (const char*)user_info[k];
    k_p2p_append<<<pp_div_up(n_user, kBlock), kBlock, 0, s>>>(x.n_in, n_user, user_elems, ut, ld, d_dst, in_elems);
  }
  PP_KERNEL_CHECK();
  PP_CUDA(cudaMemcpyAsync(w.host_stats, w.dev_stats, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaEventRecord(w.ev, s));
  std::vector<const void*> info(ps->nmembers, nullptr);
  for (int k = 0; k < ps->nmembers; ++k) info[k] = in_data[k];
  pp_status st = pp_ps_rebuild_ex(ps, new_element, (int32_t)ld, x.n_total, ld, in_elems, info.data(), (pp_stream)s);
  for (char* p : in_data) pp_dev_free(p, s);
  pp_dev_free(in_elems, s); pp_dev_free(d_dst, s);
  PP_CUDA(cudaEventSynchronize(w.ev));       // long done: the rebuild read its scalars after this copy
  if (stats_host) {
    stats_host->sent = w.host_stats[0]; stats_host->received = w.host_stats[1];
    stats_host->deferred = w.host_stats[2];
  }
  if (w.host_stats[3]) {
    pp_set_error("migrate: a peer did not publish its particles within 20 s");
    return PP_ERR_NCCL;
  }
  return st;
}
}  // namespace

extern "C" void pp_comm_set_p2p(int32_t enable) { g_p2p_enable = enable ? 1 : 0; }
extern "C" pp_status pp_comm_set_p2p_window(pp_comm* c, int64_t bytes_per_peer) {
  PP_REQUIRE(c && bytes_per_peer >= 4096, "bad argument");
  PP_REQUIRE(!c->p2p.tried, "the peer-memory window is sized before the first migration");
  c->p2p.seg_bytes = (size_t)bytes_per_peer & ~(size_t)15;
  c->p2p.seg_set = true;
  return PP_OK;
}
extern "C" int32_t pp_comm_p2p_active(const pp_comm* c) { return c && c->p2p.ok ? 1 : 0; }

extern "C" pp_status pp_ps_migrate(pp_ps* ps, pp_comm* comm, int32_t* new_element,
                                   const int32_t* new_process, int32_t n_new,
                                   const int32_t* new_particle_elements,
                                   const void* const* new_particle_info,
                                   pp_migrate_stats* stats_host, pp_stream stream_) {
  PP_REQUIRE(ps && comm && (new_element || ps->capacity == 0), "null argument");
  cudaStream_t s = (cudaStream_t)stream_;
  PP_TIME_KIND(s, ps->cfg.kind, "particle migration");   // SCS_migrate.h:218 (here including the rebuild)
  if (stats_host) { stats_host->sent = 0; stats_host->received = 0; stats_host->deferred = 0; }
  // serial: SCS_migrate.h:20-25
  if (comm->nranks == 1)
    return pp_ps_rebuild(ps, new_element, n_new, new_particle_elements, new_particle_info, stream_);
  PP_REQUIRE(new_process || ps->capacity == 0, "new_process is required");
  PP_REQUIRE(ps->nmembers <= 16, "at most 16 particle members are supported");
  const int R = comm->nranks, me = comm->rank;
  PackTable pt;
  pt.n = ps->nmembers;
  for (int k = 0; k < pt.n; ++k) {
    pt.src[k] = (const char*)ps->data[k];
    pt.bytes[k] = ps->members[k].scalar_bytes;
    pt.ncomp[k] = ps->members[k].ncomp;
  }
  // peer-memory window (set up collectively by the first migration of this communicator)
  // default segment: room for 1/16 of this structure's slots per peer, at least 24 MiB
  if (!comm->p2p.tried) PP_TRY(p2p_setup(comm, (size_t)ps->capacity / 16 * p2p_rec_bytes(pt), s));
  if (comm->p2p.ok) {
    bool plain = true;                   // members in 1/2/4/8-byte scalars: always
    if (plain)
      return migrate_p2p(ps, comm, new_element, new_process, n_new, new_particle_elements, new_particle_info,
                         stats_host, pt, s);
  }
  // ---- NCCL path (no peer access between the GPUs, or pp_comm_set_p2p(0))
  PP_NEED_NCCL(comm);
  DevScope scratch(s);                 // every early return below frees what was allocated so far
  // 1. particles per destination, 2. counts to everybody (PS_Comm_Ialltoall, :48)
  int *send_cnt, *all_cnt;
  PP_TRY(scratch.alloc(&send_cnt, R));
  PP_TRY(scratch.alloc(&all_cnt, (size_t)R * R));
  PP_CUDA(cudaMemsetAsync(send_cnt, 0, sizeof(int) * R, s));
  if (ps->capacity > 0)
    k_count_dest<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(ps->view(), new_process, new_element, me, R,
                                                                    send_cnt);
  PP_KERNEL_CHECK();
  PP_NCCL(g_nccl.AllGather(send_cnt, all_cnt, (size_t)R, ncclInt32, comm->comm, s));
  std::vector<int> h_all((size_t)R * R);
  PP_CUDA(cudaMemcpyAsync(h_all.data(), all_cnt, sizeof(int) * R * R, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  std::vector<int> h_send(R), h_recv(R), h_recv_off(R + 1, 0);
  std::vector<size_t> h_sbo(R + 1, 0), h_rbo(R + 1, 0);
  long tot_send = 0, tot_recv = 0;
  for (int p = 0; p < R; ++p) {
    h_send[p] = h_all[(size_t)me * R + p];
    h_recv[p] = h_all[(size_t)p * R + me];
    h_sbo[p + 1] = h_sbo[p] + block_bytes(pt, (size_t)h_send[p]);
    h_rbo[p + 1] = h_rbo[p] + block_bytes(pt, (size_t)h_recv[p]);
    h_recv_off[p + 1] = h_recv_off[p] + h_recv[p];
    tot_send += h_send[p]; tot_recv += h_recv[p];
  }
  if (stats_host) { stats_host->sent = tot_send; stats_host->received = tot_recv; }
  // 3. pack: one buffer per peer holding gids + every member (gatherParticlesToSend +
  //    CopyParticlesToSend, :83-98), sent particles are marked deleted in new_element
  char *sendbuf, *recvbuf;
  size_t *d_sbo, *d_rbo;
  int *cursor, *d_recv_cnt, *d_recv_off;
  PP_TRY(scratch.alloc(&sendbuf, h_sbo[R] + 8));
  PP_TRY(scratch.alloc(&recvbuf, h_rbo[R] + 8));
  PP_TRY(scratch.alloc(&d_sbo, R + 1));
  PP_TRY(scratch.alloc(&d_rbo, R + 1));
  PP_TRY(scratch.alloc(&cursor, R));
  PP_TRY(scratch.alloc(&d_recv_cnt, R));
  PP_TRY(scratch.alloc(&d_recv_off, R + 1));
  PP_CUDA(cudaMemcpyAsync(d_sbo, h_sbo.data(), sizeof(size_t) * (R + 1), cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemcpyAsync(d_rbo, h_rbo.data(), sizeof(size_t) * (R + 1), cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemcpyAsync(d_recv_cnt, h_recv.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemcpyAsync(d_recv_off, h_recv_off.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice, s));
  PP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * R, s));
  if (ps->capacity > 0)
    k_pack<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(
        ps->view(), new_process, new_element, (const long long*)ps->elem_gids, me, R, send_cnt, d_sbo,
        cursor, pt, ps->stride, sendbuf);
  PP_KERNEL_CHECK();
  // 4. all-to-all-v: grouped send/recv, one message per peer (:148-175 uses 1+num_types per peer)
  {
    NcclGroup group;                   // closes the group on every path: an open group would hang the peers
    PP_TRY(group.begin());
    for (int p = 0; p < R; ++p) {
      if (p == me) continue;
      if (h_send[p] > 0)
        PP_NCCL(g_nccl.Send(sendbuf + h_sbo[p], h_sbo[p + 1] - h_sbo[p], ncclUint8, p, comm->comm, s));
      if (h_recv[p] > 0)
        PP_NCCL(g_nccl.Recv(recvbuf + h_rbo[p], h_rbo[p + 1] - h_rbo[p], ncclUint8, p, comm->comm, s));
    }
    PP_TRY(group.end());
  }
  // 5. unpack into contiguous new-particle arrays, gid -> lid, append the caller's new particles
  const int n_in = (int)tot_recv + n_new;
  std::vector<char*> in_data(ps->nmembers, nullptr);
  int* in_elems = nullptr;
  int* bad;
  PP_TRY(scratch.alloc(&bad, 1));
  PP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
  pp_status st = PP_OK;
  if (n_in > 0) {
    PP_TRY(scratch.alloc(&in_elems, n_in));
    for (int k = 0; k < ps->nmembers; ++k)
      PP_TRY(scratch.alloc(&in_data[k], (size_t)pt.bytes[k] * pt.ncomp[k] * n_in));
    // [ncomp][n_in] with received particles first; the caller's new particles follow
    if (tot_recv > 0) {
      // received particles are written with row length n_in so both groups share one array
      char** d_dst;
      PP_TRY(scratch.alloc(&d_dst, ps->nmembers));
      PP_CUDA(cudaMemcpyAsync(d_dst, in_data.data(), sizeof(char*) * ps->nmembers, cudaMemcpyHostToDevice, s));
      if (ps->elem_gids && !ps->sorted_gid) {
        // lazily build the sorted gid table (createGlobalMapping, SCS_buildFns.h:101-112)
        long long* keys_in = (long long*)ps->elem_gids;
        int* vals_in;
        PP_TRY(scratch.alloc(&vals_in, ps->nelems));
        PP_TRY(pp_dev_alloc(&ps->sorted_gid, ps->nelems, s));
        PP_TRY(pp_dev_alloc(&ps->sorted_lid, ps->nelems, s));
        k_iota<<<pp_div_up(ps->nelems, kBlock), kBlock, 0, s>>>(vals_in, ps->nelems);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, (long long*)ps->sorted_gid, vals_in,
                                        ps->sorted_lid, ps->nelems, 0, 64, s);
        char* tmp;
        PP_TRY(scratch.alloc(&tmp, tb));
        PP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, (long long*)ps->sorted_gid, vals_in,
                                                ps->sorted_lid, ps->nelems, 0, 64, s));
      }
      // rows are n_in long so that the caller's new particles can follow the received ones
      k_unpack<<<pp_div_up(tot_recv, kBlock), kBlock, 0, s>>>(
          recvbuf, d_rbo, d_recv_cnt, d_recv_off, R, (int)tot_recv, n_in, pt, d_dst,
          (const long long*)ps->sorted_gid, ps->sorted_lid, ps->nelems, in_elems, bad);
    }
    if (n_new > 0) {
      PP_CUDA(cudaMemcpyAsync(in_elems + tot_recv, new_particle_elements, sizeof(int) * n_new,
                              cudaMemcpyDeviceToDevice, s));
      for (int k = 0; k < ps->nmembers; ++k) {
        const size_t sb = pt.bytes[k];
        for (int c = 0; c < pt.ncomp[k]; ++c)
          PP_CUDA(cudaMemcpyAsync(in_data[k] + ((size_t)c * n_in + tot_recv) * sb,
                                  (const char*)new_particle_info[k] + (size_t)c * n_new * sb,
                                  (size_t)n_new * sb, cudaMemcpyDeviceToDevice, s));
      }
    }
    PP_KERNEL_CHECK();
    int h_bad = 0;
    PP_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
    if (h_bad) {
      pp_set_error("migrate: a received particle names an element gid that is not on this rank");
      st = PP_ERR_INVALID;
    }
  }
  // 6. rebuild with the received particles as new particles (:209)
  if (st == PP_OK) {
    std::vector<const void*> info(ps->nmembers, nullptr);
    for (int k = 0; k < ps->nmembers; ++k) info[k] = in_data[k];
    st = pp_ps_rebuild(ps, new_element, n_in, in_elems, n_in > 0 ? info.data() : nullptr, stream_);
  }
  return st;
}

// ------------------------------------------------------------------------------------------
// Mesh::reduceCommArray for partially buffered PICparts (pumipic_comm.cpp:249-439) and the
// set-up half of Mesh::setupComm (:12-184).
//
// The reference renumbers every entity into a "bulk communication ordering" (whole cores first,
// boundary entities by an atomic counter), stages the array through the host and exchanges one
// MPI message per buffered core plus one per bounding part, merging with device atomics.  Here
// the plan is a plain owner fan-in / fan-out over explicit index lists: for every peer p the
// entities this rank holds that p owns (send list) and the entities this rank owns that p holds
// (receive list), matched once by global id at set-up.  A reduction is pack -> grouped
// ncclSend/ncclRecv over NVLink -> merge in ascending rank order (deterministic, no atomics) ->
// pack -> send back -> unpack; nothing touches the host.
// ------------------------------------------------------------------------------------------
struct pp_comm_plan {
  pp_comm* comm;
  int64_t nents;
  std::vector<int64_t> send_off, recv_off;   // [R+1], entity offsets per peer
  int* d_send_idx;                           // local index of every entity to send, grouped by owner
  int* d_recv_idx;                           // local index of every entity received, grouped by holder
};

namespace {
template <class T>
__global__ void k_plan_pack(const T* __restrict__ arr, const int* __restrict__ idx, long n, int nvals,
                            T* __restrict__ buf) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n * nvals) return;
  buf[i] = arr[(long)idx[i / nvals] * nvals + i % nvals];
}
template <class T>
__global__ void k_plan_unpack(T* __restrict__ arr, const int* __restrict__ idx, long n, int nvals,
                              const T* __restrict__ buf) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n * nvals) return;
  arr[(long)idx[i / nvals] * nvals + i % nvals] = buf[i];
}
template <class T>
__global__ void k_plan_merge(T* __restrict__ arr, const int* __restrict__ idx, long n, int nvals,
                             const T* __restrict__ buf, int op) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n * nvals) return;
  T* a = arr + (long)idx[i / nvals] * nvals + i % nvals;
  const T x = *a, y = buf[i];
  *a = op == PP_SUM ? x + y : op == PP_MAX ? (x < y ? y : x) : (y < x ? y : x);
}

template <class T>
pp_status plan_reduce_t(pp_comm_plan* pl, T* arr, int nvals, int op, ncclDataType_t nt, cudaStream_t s) {
  pp_comm* c = pl->comm;
  const int R = c->nranks, me = c->rank;
  const int64_t ns = pl->send_off[R], nr = pl->recv_off[R];
  T *sbuf, *rbuf;
  PP_TRY(pp_dev_alloc(&sbuf, (size_t)ns * nvals, s));
  PP_TRY(pp_dev_alloc(&rbuf, (size_t)nr * nvals, s));
  auto blocks = [&](int64_t n) { return pp_div_up(n * nvals, kBlock); };
  if (op != PP_BCAST) {
    // fan in: every copy goes to the owner of its entity
    if (ns) k_plan_pack<<<blocks(ns), kBlock, 0, s>>>(arr, pl->d_send_idx, ns, nvals, sbuf);
    PP_KERNEL_CHECK();
    NcclGroup group3;
    PP_TRY(group3.begin());
    for (int p = 0; p < R; ++p) {
      if (p == me) continue;
      const int64_t a = pl->send_off[p], b = pl->send_off[p + 1];
      if (b > a) PP_NCCL(g_nccl.Send(sbuf + a * nvals, (size_t)(b - a) * nvals, nt, p, c->comm, s));
      const int64_t ra = pl->recv_off[p], rb = pl->recv_off[p + 1];
      if (rb > ra) PP_NCCL(g_nccl.Recv(rbuf + ra * nvals, (size_t)(rb - ra) * nvals, nt, p, c->comm, s));
    }
    PP_TRY(group3.end());
    for (int p = 0; p < R; ++p) {          // ascending rank order: bit-reproducible sums
      const int64_t ra = pl->recv_off[p], rb = pl->recv_off[p + 1];
      if (rb > ra)
        k_plan_merge<<<blocks(rb - ra), kBlock, 0, s>>>(arr, pl->d_recv_idx + ra, rb - ra, nvals,
                                                        rbuf + ra * nvals, op);
    }
    PP_KERNEL_CHECK();
  }
  // fan out: the owner's value (the total) goes back to every copy
  if (nr) k_plan_pack<<<blocks(nr), kBlock, 0, s>>>(arr, pl->d_recv_idx, nr, nvals, rbuf);
  PP_KERNEL_CHECK();
  NcclGroup group4;
  PP_TRY(group4.begin());
  for (int p = 0; p < R; ++p) {
    if (p == me) continue;
    const int64_t ra = pl->recv_off[p], rb = pl->recv_off[p + 1];
    if (rb > ra) PP_NCCL(g_nccl.Send(rbuf + ra * nvals, (size_t)(rb - ra) * nvals, nt, p, c->comm, s));
    const int64_t a = pl->send_off[p], b = pl->send_off[p + 1];
    if (b > a) PP_NCCL(g_nccl.Recv(sbuf + a * nvals, (size_t)(b - a) * nvals, nt, p, c->comm, s));
  }
  PP_TRY(group4.end());
  if (ns) k_plan_unpack<<<blocks(ns), kBlock, 0, s>>>(arr, pl->d_send_idx, ns, nvals, sbuf);
  PP_KERNEL_CHECK();
  pp_dev_free(sbuf, s); pp_dev_free(rbuf, s);
  return PP_OK;
}
}  // namespace

extern "C" pp_status pp_comm_plan_create(pp_comm* c, int64_t nents, const int64_t* ent_gids,
                                         const int32_t* ent_owner, int32_t memspace, pp_stream stream,
                                         pp_comm_plan** out) {
  PP_REQUIRE(c && out && nents >= 0 && (nents == 0 || (ent_gids && ent_owner)), "bad argument");
  PP_REQUIRE(nents < (int64_t)1 << 31, "too many entities");
  cudaStream_t s = (cudaStream_t)stream;
  const int R = c->nranks, me = c->rank;
  PP_REQUIRE(c->nranks == 1 || c->comm, "comm plans (owner fan-in / fan-out) use NCCL: create the communicator with NCCL");
  pp_comm_plan* pl = new pp_comm_plan();
  pl->comm = c; pl->nents = nents; pl->d_send_idx = nullptr; pl->d_recv_idx = nullptr;
  pl->send_off.assign(R + 1, 0); pl->recv_off.assign(R + 1, 0);
  *out = pl;
  if (R == 1) return PP_OK;
  std::vector<int64_t> gid((size_t)nents);
  std::vector<int32_t> own((size_t)nents);
  if (nents) {
    const cudaMemcpyKind k = memspace == PP_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost;
    PP_CUDA(cudaMemcpyAsync(gid.data(), ent_gids, sizeof(int64_t) * nents, k, s));
    PP_CUDA(cudaMemcpyAsync(own.data(), ent_owner, sizeof(int32_t) * nents, k, s));
    PP_CUDA(cudaStreamSynchronize(s));
  }
  // send lists: my copies of entities owned elsewhere, grouped by owner, ascending gid
  std::vector<std::vector<std::pair<int64_t, int>>> bucket(R);
  std::vector<std::pair<int64_t, int>> mine;
  for (int64_t i = 0; i < nents; ++i) {
    PP_REQUIRE(own[i] >= 0 && own[i] < R, "entity owner out of range");
    (own[i] == me ? mine : bucket[own[i]]).push_back({gid[i], (int)i});
  }
  std::sort(mine.begin(), mine.end());
  std::vector<int> send_idx, h_cnt(R, 0);
  std::vector<int64_t> send_gid;
  for (int p = 0; p < R; ++p) {
    std::sort(bucket[p].begin(), bucket[p].end());
    h_cnt[p] = (int)bucket[p].size();
    pl->send_off[p + 1] = pl->send_off[p] + h_cnt[p];
    for (auto& e : bucket[p]) { send_gid.push_back(e.first); send_idx.push_back(e.second); }
  }
  // counts to everybody
  int *d_cnt, *d_all;
  PP_TRY(pp_dev_alloc(&d_cnt, R, s));
  PP_TRY(pp_dev_alloc(&d_all, (size_t)R * R, s));
  PP_CUDA(cudaMemcpyAsync(d_cnt, h_cnt.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s));
  PP_NCCL(g_nccl.AllGather(d_cnt, d_all, (size_t)R, ncclInt32, c->comm, s));
  std::vector<int> h_all((size_t)R * R);
  PP_CUDA(cudaMemcpyAsync(h_all.data(), d_all, sizeof(int) * R * R, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  for (int p = 0; p < R; ++p) pl->recv_off[p + 1] = pl->recv_off[p] + h_all[(size_t)p * R + me];
  const int64_t ns = pl->send_off[R], nr = pl->recv_off[R];
  // global ids of what every holder will send me
  long long *d_sg, *d_rg;
  PP_TRY(pp_dev_alloc(&d_sg, (size_t)ns, s));
  PP_TRY(pp_dev_alloc(&d_rg, (size_t)nr, s));
  if (ns) PP_CUDA(cudaMemcpyAsync(d_sg, send_gid.data(), sizeof(int64_t) * ns, cudaMemcpyHostToDevice, s));
  NcclGroup group5;
  PP_TRY(group5.begin());
  for (int p = 0; p < R; ++p) {
    if (p == me) continue;
    const int64_t a = pl->send_off[p], b = pl->send_off[p + 1];
    if (b > a) PP_NCCL(g_nccl.Send(d_sg + a, (size_t)(b - a), ncclInt64, p, c->comm, s));
    const int64_t ra = pl->recv_off[p], rb = pl->recv_off[p + 1];
    if (rb > ra) PP_NCCL(g_nccl.Recv(d_rg + ra, (size_t)(rb - ra), ncclInt64, p, c->comm, s));
  }
  PP_TRY(group5.end());
  std::vector<int64_t> recv_gid((size_t)nr);
  if (nr) PP_CUDA(cudaMemcpyAsync(recv_gid.data(), d_rg, sizeof(int64_t) * nr, cudaMemcpyDeviceToHost, s));
  PP_CUDA(cudaStreamSynchronize(s));
  pp_dev_free(d_cnt, s); pp_dev_free(d_all, s); pp_dev_free(d_sg, s); pp_dev_free(d_rg, s);
  std::vector<int> recv_idx((size_t)nr);
  for (int64_t k = 0; k < nr; ++k) {
    auto it = std::lower_bound(mine.begin(), mine.end(), std::make_pair(recv_gid[k], -1));
    if (it == mine.end() || it->first != recv_gid[k]) {
      pp_set_error("pp_comm_plan_create: rank %d received entity gid %lld that it does not own", me,
                   (long long)recv_gid[k]);
      return PP_ERR_INVALID;
    }
    recv_idx[k] = it->second;
  }
  PP_CUDA(cudaMalloc(&pl->d_send_idx, sizeof(int) * (size_t)(ns ? ns : 1)));
  PP_CUDA(cudaMalloc(&pl->d_recv_idx, sizeof(int) * (size_t)(nr ? nr : 1)));
  if (ns) PP_CUDA(cudaMemcpy(pl->d_send_idx, send_idx.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
  if (nr) PP_CUDA(cudaMemcpy(pl->d_recv_idx, recv_idx.data(), sizeof(int) * nr, cudaMemcpyHostToDevice));
  return PP_OK;
}

extern "C" pp_status pp_comm_plan_destroy(pp_comm_plan* pl) {
  if (!pl) return PP_OK;
  if (pl->d_send_idx) cudaFree(pl->d_send_idx);
  if (pl->d_recv_idx) cudaFree(pl->d_recv_idx);
  delete pl;
  return PP_OK;
}

extern "C" pp_status pp_comm_plan_counts(const pp_comm_plan* pl, int64_t* n_send, int64_t* n_recv) {
  PP_REQUIRE(pl, "null plan");
  if (n_send) *n_send = pl->send_off.back();
  if (n_recv) *n_recv = pl->recv_off.back();
  return PP_OK;
}

extern "C" pp_status pp_comm_plan_reduce(pp_comm_plan* pl, void* comm_array, int32_t nvals, int32_t dtype,
                                         int32_t op, pp_stream stream) {
  PP_REQUIRE(pl && (comm_array || pl->nents == 0) && nvals >= 1, "bad argument");
  PP_REQUIRE(op == PP_SUM || op == PP_MAX || op == PP_MIN || op == PP_BCAST, "unknown reduction");
  if (pl->comm->nranks == 1) return PP_OK;     // pumipic_comm.cpp:232-233
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case PP_INT32: return plan_reduce_t(pl, (int*)comm_array, nvals, op, ncclInt32, s);
    case PP_INT64: return plan_reduce_t(pl, (long long*)comm_array, nvals, op, ncclInt64, s);
    case PP_FLOAT32: return plan_reduce_t(pl, (float*)comm_array, nvals, op, ncclFloat32, s);
    case PP_FLOAT64: return plan_reduce_t(pl, (double*)comm_array, nvals, op, ncclFloat64, s);
    default: pp_set_error("unknown pp_dtype %d", dtype); return PP_ERR_INVALID;
  }
}
