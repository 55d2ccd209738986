// pumipic_b200.hpp -- header-only C++ mirror of the PUMI-PIC API surface of the particle hot path,
// implemented over the C ABI of libpumipic_b200.so (include/pumipic_b200.h).
//
// Same names, argument meaning and error behaviour as the reference so that a driver written
// against SCOREC/pumi-pic (test/pseudoPushAndSearch.cpp, test/test_adj.cpp, test/pseudoXGCm.cpp)
// reads the same here.  What it mirrors (paths relative to the reference):
//   pumipic::MemberTypes / BaseType          particle_structs/src/support/MemberTypes.h:21-60,
//                                            support/ppTypes.h:13-30
//   pumipic::Segment                         particle_structs/src/support/Segment.h:31-101
//   pumipic::ParticleStructure<DataTypes>    particle_structs/src/particle_structure.hpp:19-144
//   pumipic::SellCSigma / SCS_Input          scs/SellCSigma.h:66-72, scs/scs_input.hpp:4-36
//   pumipic::CSR / DPS / CabM                csr/CSR.hpp:37-44, dps/dps.hpp:41-48, cabm/cabm.hpp:41-48
//   pumipic::parallel_for                    particle_structs/src/ps_for.hpp:5-31
//   pumipic::Mesh                            src/pumipic_mesh.hpp:12-76 (hot-path subset)
//   pumipic::search_mesh / search_mesh_2d    src/pumipic_adjacency.hpp:37-45,559-562,1013-1020
//   pumipic::migrate_ptcls / setUnsafeProcs  src/pumipic_ptcl_ops.hpp:12-85
//   pumipic::Distributor                     particle_structs/src/support/psDistributor.hpp:10-28
//   pumipic::ParticleBalancer / migrate_lb_ptcls   src/pumipic_lb.hpp:32-115, pumipic_ptcl_ops.hpp:55-71
//   PS_Comm_* (ViewComm)                     support/ViewComm.h:51-291
//   getPIDs / getMemberView                  particle_structs/src/ps_for.hpp:57-88, MemberTypeLibraries.h:90-105
//
// Differences forced by the missing third-party stack (no Kokkos, no Omega_h here): device arrays
// are pumipic::View<T> (a ref-counted cudaMalloc buffer with data()/size(), the role of
// Kokkos::View<T*> and Omega_h::Write<T>), and pumipic::Mesh is built from the mesh arrays an
// Omega_h mesh would hand over.  Errors: like the reference (printError + exit / throw 1), a failed
// C-ABI call prints pp_last_error() and throws std::runtime_error.
//
// User lambdas cannot cross a C ABI: ps::parallel_for is a template that needs nvcc
// (__CUDACC__); host-only translation units get everything else, including the named built-in
// kernels (push, search, rebuild, migrate, scatter).
#pragma once

#include <cuda_runtime.h>

#include <cctype>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "pumipic_b200.h"

#ifdef __CUDACC__
#define PP_INLINE __host__ __device__ inline
#define PP_DEV_INLINE __device__ inline
#define PS_LAMBDA [=] __device__
#else
#define PP_INLINE inline
#define PP_DEV_INLINE inline
#endif

namespace pumipic {

typedef int lid_t;       // support/ppTypes.h:6-8
typedef long int gid_t;
typedef double fp_t;     // src/pumipic_kktypes.hpp:10-16 (FP64 build)

inline void pp_check(pp_status st, const char* what) {
  if (st != PP_OK) {
    std::fprintf(stderr, "[ERROR] %s: %s\n", what, pp_last_error());
    throw std::runtime_error(std::string(what) + ": " + pp_last_error());
  }
}
inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[ERROR] %s: %s\n", what, cudaGetErrorString(e));
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
}

// ---------------------------------------------------------------- timing (support/ppTiming.hpp:28-75)
// The library's phases record themselves under the reference's labels; these are the calls an
// application already makes.
enum TimingSortOption { SORT_ALPHA, SORT_ORDER, SORT_LONGEST, SORT_SHORTEST };
inline void SetTimingVerbosity(int verbosity) { pp_timing_set_verbosity(verbosity); }
inline void EnableTiming() { pp_timing_enable(1); }
inline void DisableTiming() { pp_timing_enable(0); }
inline void RecordTime(const std::string& str, double seconds, double /*prebarrierTime*/ = 0.0) {
  pp_timing_record(str.c_str(), seconds);
}
inline void SummarizeTime(TimingSortOption sort = SORT_ALPHA) { pp_timing_summarize((int)sort); }

// ---------------------------------------------------------------- device array
template <class T>
class View {
 public:
  View() : n_(0) {}
  explicit View(size_t n, const std::string& name = "") : n_(n), name_(name) { alloc(); }
  View(const std::string& name, size_t n) : n_(n), name_(name) { alloc(); }   // Kokkos argument order
  View(size_t n, T fill, const std::string& name = "") : n_(n), name_(name) {  // Omega_h::Write(n, v)
    alloc();
    std::vector<T> h(n, fill);
    if (n) cuda_check(cudaMemcpy(p_.get(), h.data(), n * sizeof(T), cudaMemcpyHostToDevice), "View fill");
  }
  explicit View(const std::vector<T>& host) : n_(host.size()) {
    alloc();
    if (n_) cuda_check(cudaMemcpy(p_.get(), host.data(), n_ * sizeof(T), cudaMemcpyHostToDevice), "View h2d");
  }
  T* data() const { return p_.get(); }
  size_t size() const { return n_; }
  bool exists() const { return n_ != 0; }
  PP_DEV_INLINE T& operator()(size_t i) const { return raw_[i]; }
  PP_DEV_INLINE T& operator[](size_t i) const { return raw_[i]; }
  std::vector<T> toHost() const {
    std::vector<T> h(n_);
    if (n_) cuda_check(cudaMemcpy(h.data(), p_.get(), n_ * sizeof(T), cudaMemcpyDeviceToHost), "View d2h");
    return h;
  }

 private:
  void alloc() {
    T* q = nullptr;
    if (n_) {
      cuda_check(cudaMalloc((void**)&q, n_ * sizeof(T)), "View alloc");
      cuda_check(cudaMemset(q, 0, n_ * sizeof(T)), "View zero-fill");   // Kokkos::View zero-initialises
    }
    p_ = std::shared_ptr<T>(q, [](T* x) { if (x) cudaFree(x); });
    raw_ = q;
  }
  std::shared_ptr<T> p_;
  T* raw_ = nullptr;
  size_t n_;
  std::string name_;
};

// ---------------------------------------------------------------- member types
template <class T> struct BaseType { typedef T type; static constexpr int size = 1; };
template <class T, int N> struct BaseType<T[N]> {
  typedef typename BaseType<T>::type type;
  static constexpr int size = N * BaseType<T>::size;
};

template <typename... Types> struct MemberTypes;
template <> struct MemberTypes<> {
  static constexpr std::size_t size = 0;
  static constexpr std::size_t memsize = 0;
  static void describe(pp_member_desc*) {}
};
template <typename H, typename... T> struct MemberTypes<H, T...> {
  static constexpr std::size_t size = 1 + MemberTypes<T...>::size;
  static constexpr std::size_t memsize = sizeof(H) + MemberTypes<T...>::memsize;
  static void describe(pp_member_desc* d) {
    d->scalar_bytes = (int32_t)sizeof(typename BaseType<H>::type);
    d->ncomp = BaseType<H>::size;
    MemberTypes<T...>::describe(d + 1);
  }
};
template <std::size_t N, typename... Types> struct MemberTypeAtIndex;
template <typename H, typename... T> struct MemberTypeAtIndex<0, MemberTypes<H, T...>> { typedef H type; };
template <std::size_t N, typename H, typename... T> struct MemberTypeAtIndex<N, MemberTypes<H, T...>> {
  typedef typename MemberTypeAtIndex<N - 1, MemberTypes<T...>>::type type;
};

// MemberTypeViews (MemberTypeLibraries.h:17): one device array per member, [ncomp][n]; the entry
// after the last member keeps n so that getMemberView can index component-major
typedef void** MemberTypeViews;
template <class DataTypes> MemberTypeViews createMemberViews(int n) {
  pp_member_desc d[DataTypes::size ? DataTypes::size : 1];
  DataTypes::describe(d);
  void** v = new void*[DataTypes::size + 1];
  for (std::size_t i = 0; i < DataTypes::size; ++i)
    cuda_check(cudaMalloc(&v[i], (size_t)d[i].scalar_bytes * d[i].ncomp * (n > 0 ? n : 1)), "createMemberViews");
  v[DataTypes::size] = reinterpret_cast<void*>(static_cast<std::intptr_t>(n > 0 ? n : 0));
  return v;
}
template <class DataTypes> void destroyViews(MemberTypeViews v) {
  if (!v) return;
  for (std::size_t i = 0; i < DataTypes::size; ++i) cudaFree(v[i]);
  delete[] v;
}
// getMemberView<DataTypes, N>(mtv) (MemberTypeLibraries.h:90-105): typed access to the N-th array
// of a MemberTypeViews; view(i) / view(i, c) = particle i, component c (device code), plus
// host <-> device copies of the whole array in the same [ncomp][n] order.
template <class Type>
class MemberView {
 public:
  typedef typename BaseType<Type>::type Base;
  static constexpr int ncomp = BaseType<Type>::size;
  MemberView() : p_(nullptr), n_(0) {}
  MemberView(Base* p, long n) : p_(p), n_(n) {}
  PP_DEV_INLINE Base& operator()(const int& i) const { return p_[i]; }
  PP_DEV_INLINE Base& operator()(const int& i, const int& c) const { return p_[(long)c * n_ + i]; }
  Base* data() const { return p_; }
  long size() const { return n_; }
  void fromHost(const std::vector<Base>& h) const {      // h[c * n + i]
    if (n_) cuda_check(cudaMemcpy(p_, h.data(), sizeof(Base) * ncomp * n_, cudaMemcpyHostToDevice), "MemberView h2d");
  }
  std::vector<Base> toHost() const {
    std::vector<Base> h((size_t)ncomp * n_);
    if (n_) cuda_check(cudaMemcpy(h.data(), p_, sizeof(Base) * ncomp * n_, cudaMemcpyDeviceToHost), "MemberView d2h");
    return h;
  }

 private:
  Base* p_;
  long n_;
};
template <class DataTypes, std::size_t N>
MemberView<typename MemberTypeAtIndex<N, DataTypes>::type> getMemberView(MemberTypeViews v) {
  typedef typename MemberTypeAtIndex<N, DataTypes>::type T;
  const long n = (long)reinterpret_cast<std::intptr_t>(v[DataTypes::size]);
  return MemberView<T>(static_cast<typename BaseType<T>::type*>(v[N]), n);
}

// ---------------------------------------------------------------- Segment: by-value accessor
// value(slot, i) lives at base[i * stride + slot] (LayoutLeft, support/ppView.h:7)
template <class Type>
class Segment {
 public:
  typedef typename BaseType<Type>::type Base;
  Segment() : base_(nullptr), stride_(0) {}
  Segment(Base* b, long stride) : base_(b), stride_(stride) {}
  PP_DEV_INLINE Base& operator()(const int& particle_index) const { return base_[particle_index]; }
  PP_DEV_INLINE Base& operator()(const int& particle_index, const int& i) const {
    return base_[(long)i * stride_ + particle_index];
  }
  Base* data() const { return base_; }
  long stride() const { return stride_; }

 private:
  Base* base_;
  long stride_;
};

// ---------------------------------------------------------------- Distributor (world only)
class Distributor {
 public:
  Distributor() : comm_(nullptr) {}
  explicit Distributor(pp_comm* c) : comm_(c) {}
  bool isWorld() const { return true; }
  // psDistributor.hpp:24-28: position of a rank in the distributor and back; the world
  // distributor is the identity (usable in device code)
  PP_INLINE int index(int rank) const { return rank; }
  PP_INLINE int rank(int i) const { return i; }
  int num_ranks() const { return comm_ ? pp_comm_size(comm_) : 1; }
  pp_comm* comm() const { return comm_; }

 private:
  pp_comm* comm_;
};

// Kokkos::TeamPolicy stand-in: only team_size() matters to the structures (C = team_size)
struct TeamPolicy {
  TeamPolicy(int league = 1, int team = 32) : league_(league), team_(team) {}
  int team_size() const { return team_; }
  int league_, team_;
};

enum PaddingStrategy { PAD_EVENLY, PAD_PROPORTIONALLY, PAD_INVERSELY };

// ---------------------------------------------------------------- ParticleStructure
template <class DataTypes>
class ParticleStructure {
 public:
  typedef DataTypes Types;
  typedef View<lid_t> kkLidView;
  typedef View<gid_t> kkGidView;
  typedef MemberTypeViews MTVs;
  template <std::size_t N> using DataType = typename MemberTypeAtIndex<N, DataTypes>::type;
  template <std::size_t N> using Slice = Segment<DataType<N>>;

  virtual ~ParticleStructure() { if (h_) pp_ps_destroy(h_); }
  const std::string& getName() const { return name; }
  lid_t nElems() const { return pp_ps_nelems(h_); }
  lid_t nPtcls() const { return pp_ps_nptcls(h_); }
  lid_t capacity() const { return pp_ps_capacity(h_); }
  lid_t numRows() const { return pp_ps_numrows(h_); }

  // invalidated by rebuild / migrate, like the reference's Segments
  template <std::size_t N> Slice<N> get() {
    void* base = nullptr;
    int64_t stride = 0;
    pp_check(pp_ps_member(h_, (int32_t)N, &base, &stride), "ParticleStructure::get");
    return Slice<N>(static_cast<typename BaseType<DataType<N>>::type*>(base), (long)stride);
  }

  virtual void rebuild(kkLidView new_element, kkLidView new_particle_elements = kkLidView(),
                       MTVs new_particle_info = NULL) {
    pp_check(pp_ps_rebuild(h_, new_element.data(), (int32_t)new_particle_elements.size(),
                           new_particle_elements.data(), new_particle_info, stream_),
             "ParticleStructure::rebuild");
  }
  virtual void migrate(kkLidView new_element, kkLidView new_process, Distributor dist = Distributor(),
                       kkLidView new_particle_elements = kkLidView(), MTVs new_particle_info = NULL) {
    if (!dist.comm() || dist.num_ranks() == 1) {   // SCS_migrate.h:20-25
      rebuild(new_element, new_particle_elements, new_particle_info);
      return;
    }
    pp_migrate_stats st;
    pp_check(pp_ps_migrate(h_, dist.comm(), new_element.data(), new_process.data(),
                           (int32_t)new_particle_elements.size(), new_particle_elements.data(),
                           new_particle_info, &st, stream_),
             "ParticleStructure::migrate");
  }
  // ps_for.hpp:57-88: slots of the masked particles grouped by element + each group's start
  template <typename ViewT>
  void getPIDs(ViewT& pids, ViewT& offsets) {
    offsets = ViewT((size_t)nElems() + 1);
    pids = ViewT((size_t)nPtcls());
    pp_check(pp_ps_get_pids(h_, pids.data(), offsets.data(), stream_), "ParticleStructure::getPIDs");
    cuda_check(cudaStreamSynchronize((cudaStream_t)stream_), "ParticleStructure::getPIDs");
  }
  virtual void printMetrics() const {
    std::printf("%s: elements %d rows %d particles %d capacity %d\n", name.c_str(), nElems(), numRows(),
                nPtcls(), capacity());
  }
  virtual void printFormat(const char* prefix = "") const { std::printf("%s%s\n", prefix, name.c_str()); }

  pp_ps* handle() const { return h_; }
  cudaStream_t stream() const { return (cudaStream_t)stream_; }
  void setStream(cudaStream_t s) { stream_ = (pp_stream)s; }

 protected:
  ParticleStructure(const std::string& n) : name(n), h_(nullptr), stream_(nullptr) {}
  void create(pp_ps_config cfg, lid_t ne, lid_t np, kkLidView ppe, kkGidView gids,
              kkLidView particle_elements, MTVs particle_info) {
    pp_member_desc d[DataTypes::size ? DataTypes::size : 1];
    DataTypes::describe(d);
    pp_check(pp_ps_create(&cfg, (int32_t)DataTypes::size, d, ne, np, ppe.data(),
                          gids.size() ? (const int64_t*)gids.data() : nullptr,
                          particle_elements.size() ? particle_elements.data() : nullptr,
                          particle_info, PP_DEVICE, stream_, &h_),
             name.c_str());
  }
  std::string name;
  pp_ps* h_;
  pp_stream stream_;
};

template <class DataTypes> class SellCSigma;

template <class DataTypes>
class SCS_Input {
 public:
  typedef View<lid_t> kkLidView;
  typedef View<gid_t> kkGidView;
  typedef MemberTypeViews MTVs;
  typedef TeamPolicy PolicyType;
  SCS_Input(PolicyType& p, lid_t sigma_, lid_t vertical_chunk_size, lid_t num_elements, lid_t num_particles,
            kkLidView particles_per_elements, kkGidView element_gids,
            kkLidView particle_elements_ = kkLidView(), MTVs particle_info = NULL)
      : policy(p), sig(sigma_), V(vertical_chunk_size), ne(num_elements), np(num_particles),
        ppe(particles_per_elements), e_gids(element_gids), particle_elms(particle_elements_), p_info(particle_info) {}
  bool always_realloc = false;
  double minimize_size = .8;
  double shuffle_padding = 0.1;
  double extra_padding = 0.05;
  PaddingStrategy padding_strat = PAD_EVENLY;
  std::string name = "ptcls";

 protected:
  PolicyType policy;
  lid_t sig, V, ne, np;
  kkLidView ppe;
  kkGidView e_gids;
  kkLidView particle_elms;
  MTVs p_info;
  friend class SellCSigma<DataTypes>;
};

template <class DataTypes>
class SellCSigma : public ParticleStructure<DataTypes> {
 public:
  typedef ParticleStructure<DataTypes> Base;
  using typename Base::kkGidView;
  using typename Base::kkLidView;
  using typename Base::MTVs;
  typedef TeamPolicy PolicyType;
  SellCSigma(PolicyType& p, lid_t sigma, lid_t vertical_chunk_size, lid_t num_elements, lid_t num_particles,
             kkLidView particles_per_element, kkGidView element_gids, kkLidView particle_elements = kkLidView(),
             MTVs particle_info = NULL)
      : Base("ptcls") {
    pp_ps_config cfg;
    pp_ps_config_default(&cfg, PP_PS_SCS);
    cfg.team_size = p.team_size(); cfg.sigma = sigma; cfg.V = vertical_chunk_size;
    this->create(cfg, num_elements, num_particles, particles_per_element, element_gids, particle_elements,
                 particle_info);
  }
  explicit SellCSigma(SCS_Input<DataTypes>& in) : Base(in.name) {
    pp_ps_config cfg;
    pp_ps_config_default(&cfg, PP_PS_SCS);
    cfg.team_size = in.policy.team_size(); cfg.sigma = in.sig; cfg.V = in.V;
    cfg.shuffle_padding = in.shuffle_padding; cfg.extra_padding = in.extra_padding;
    cfg.minimize_size = in.minimize_size; cfg.padding_strat = (int32_t)in.padding_strat;
    cfg.always_realloc = in.always_realloc;
    this->create(cfg, in.ne, in.np, in.ppe, in.e_gids, in.particle_elms, in.p_info);
  }
  lid_t C() const { return layout().C; }
  lid_t V() const { return layout().V; }
  pp_ps_layout layout() const {
    pp_ps_layout l;
    pp_check(pp_ps_get_layout(this->h_, this->stream_, &l), "SellCSigma::layout");
    return l;
  }
};

// The flat structures and their input classes (csr/CSR_input.hpp:10-42, dps/dps_input.hpp:9-34,
// cabm/cabm_input.hpp:9-34).  DPS and CabM are constructible from their input class (dps.hpp:48,
// cabm.hpp:48); the reference declares CSR_Input but gives CSR no constructor for it.
#define PP_B200_FLAT_INPUT(NAME, STRUCT, EXTRA_MEMBERS)                                               \
  template <class DataTypes> class STRUCT;                                                            \
  template <class DataTypes>                                                                          \
  class NAME {                                                                                        \
   public:                                                                                            \
    typedef View<lid_t> kkLidView;                                                                    \
    typedef View<gid_t> kkGidView;                                                                    \
    typedef MemberTypeViews MTVs;                                                                     \
    typedef TeamPolicy PolicyType;                                                                    \
    NAME(PolicyType& p, lid_t num_elements, lid_t num_particles, kkLidView particles_per_elements,    \
         kkGidView element_gids, kkLidView particle_elements = kkLidView(), MTVs particle_info = NULL) \
        : policy(p), ne(num_elements), np(num_particles), ppe(particles_per_elements),                \
          e_gids(element_gids), particle_elms(particle_elements), p_info(particle_info) {}            \
    EXTRA_MEMBERS                                                                                     \
    std::string name = "ptcls";                                                                       \
                                                                                                      \
   protected:                                                                                         \
    PolicyType policy;                                                                                \
    lid_t ne, np;                                                                                     \
    kkLidView ppe;                                                                                    \
    kkGidView e_gids;                                                                                 \
    kkLidView particle_elms;                                                                          \
    MTVs p_info;                                                                                      \
    friend class STRUCT<DataTypes>;                                                                   \
  };
PP_B200_FLAT_INPUT(CSR_Input, CSR, bool always_realloc = false; double minimize_size = 0.8; double padding_amount = 1.05;)
PP_B200_FLAT_INPUT(DPS_Input, DPS, double extra_padding = 0.05;)
PP_B200_FLAT_INPUT(CabM_Input, CabM, double extra_padding = 0.05;)
#undef PP_B200_FLAT_INPUT

#define PP_B200_FLAT_STRUCTURE(NAME, KIND, LABEL, INPUT_CTOR)                                         \
  template <class DataTypes>                                                                          \
  class NAME : public ParticleStructure<DataTypes> {                                                  \
   public:                                                                                            \
    typedef ParticleStructure<DataTypes> Base;                                                        \
    using typename Base::kkGidView;                                                                   \
    using typename Base::kkLidView;                                                                   \
    using typename Base::MTVs;                                                                        \
    typedef TeamPolicy PolicyType;                                                                    \
    typedef NAME##_Input<DataTypes> Input_T;                                                          \
    NAME(PolicyType& p, lid_t num_elements, lid_t num_particles, kkLidView particles_per_element,     \
         kkGidView element_gids, kkLidView particle_elements = kkLidView(), MTVs particle_info = NULL) \
        : Base(LABEL) {                                                                               \
      pp_ps_config cfg;                                                                               \
      pp_ps_config_default(&cfg, KIND);                                                               \
      cfg.team_size = p.team_size();                                                                  \
      this->create(cfg, num_elements, num_particles, particles_per_element, element_gids,             \
                   particle_elements, particle_info);                                                 \
    }                                                                                                 \
    INPUT_CTOR                                                                                        \
  };
#define PP_B200_PADDED_INPUT_CTOR(NAME, KIND)                                                         \
  explicit NAME(NAME##_Input<DataTypes>& in) : Base(in.name) {                                        \
    pp_ps_config cfg;                                                                                 \
    pp_ps_config_default(&cfg, KIND);                                                                 \
    cfg.team_size = in.policy.team_size();                                                            \
    cfg.extra_padding = in.extra_padding;                                                             \
    this->create(cfg, in.ne, in.np, in.ppe, in.e_gids, in.particle_elms, in.p_info);                  \
  }
PP_B200_FLAT_STRUCTURE(CSR, PP_PS_CSR, "ptcls", )                                            // csr/CSR.hpp:37-44
PP_B200_FLAT_STRUCTURE(DPS, PP_PS_DPS, "ptcls", PP_B200_PADDED_INPUT_CTOR(DPS, PP_PS_DPS))    // dps/dps.hpp:41-48
PP_B200_FLAT_STRUCTURE(CabM, PP_PS_CABM, "ptcls", PP_B200_PADDED_INPUT_CTOR(CabM, PP_PS_CABM)) // cabm/cabm.hpp:41-48
#undef PP_B200_PADDED_INPUT_CTOR
#undef PP_B200_FLAT_STRUCTURE

// ---------------------------------------------------------------- parallel_for (nvcc only)
#ifdef __CUDACC__
namespace detail {
template <class Fn>
__global__ void k_parallel_for(Fn fn, int capacity, const uint32_t* __restrict__ mask_bits,
                               const int* __restrict__ slot_elem) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= capacity) return;
  const bool mask = (mask_bits[slot >> 5] >> (slot & 31)) & 1u;
  fn(slot_elem[slot], slot, mask);
}
}  // namespace detail

// ps_for.hpp:5-31: fn(const int elem, const int slot, const bool mask) for every slot of the
// structure including padding.  Consecutive threads take consecutive slots, which in a
// Sell-C-sigma chunk are consecutive rows (coalesced member access), as in SellCSigma.h:528-558.
template <typename FunctionType, typename DataTypes>
void parallel_for(ParticleStructure<DataTypes>* ps, FunctionType& fn, std::string name = "") {
  (void)name;
  pp_ps_layout l;
  pp_check(pp_ps_get_layout(ps->handle(), (pp_stream)ps->stream(), &l), "parallel_for");
  if (l.capacity == 0) return;
  const int block = 256;
  detail::k_parallel_for<<<(l.capacity + block - 1) / block, block, 0, ps->stream()>>>(
      fn, l.capacity, l.mask_bits, l.slot_elem);
  cuda_check(cudaGetLastError(), "parallel_for launch");
}
#endif

// ---------------------------------------------------------------- Input (pumipic_input.hpp:8-77)
// What pumipic::Mesh(Input&) is built from: the full mesh (the host-side container that stands for
// Omega_h::Mesh on the set-up side), who owns every element, and how buffers and safe zone are made.
class Mesh;
class Input {
 public:
  enum Method { INVALID = -1, FULL, BFS, MINIMUM, NONE };   // pumipic_input.hpp:33-39
  enum Ownership { PARTITION, CLASSIFICATION };             // :42-45
  // partition file: `.ptn` one owner per element, `.cpn` one owner per class id (pumipic_input.cpp:19-89;
  // the file is read when the communicator has more than one rank, like the reference).  After a
  // `.cpn` file getPartition() holds the owners already resolved per element.
  Input(const pp_host_mesh* mesh, const char* partition_filename, Method bufferMethod_, Method safeMethod_,
        pp_comm* comm_ = nullptr)
      : m(mesh), ownership_rule(PARTITION), comm(comm_) {
    const int dim = pp_host_mesh_dim(mesh), ne = pp_host_mesh_nents(mesh, dim);
    partition.assign((size_t)ne, 0);
    if (comm && pp_comm_size(comm) > 1) {
      const std::string f(partition_filename);
      const size_t dot = f.rfind('.');
      if (dot == std::string::npos) throw std::runtime_error("Filename has no extension");
      const bool cpn = f.substr(dot + 1) == "cpn";
      pp_host_tag t;
      const int32_t* cls = nullptr;
      if (cpn && pp_host_mesh_find_tag(mesh, dim, "class_id", &t) == PP_OK) cls = static_cast<const int32_t*>(t.data);
      pp_check(pp_host_read_partition(partition_filename, ne, cls, partition.data()), "Input");
      if (cpn) ownership_rule = CLASSIFICATION;
      resolved = true;
    }
    init(bufferMethod_, safeMethod_);
  }
  Input(const pp_host_mesh* mesh, Ownership rule, const std::vector<lid_t>& partition_vector, Method bufferMethod_,
        Method safeMethod_, pp_comm* comm_ = nullptr)
      : m(mesh), ownership_rule(rule), partition(partition_vector), comm(comm_) {
    init(bufferMethod_, safeMethod_);
  }
  void printMethod() const {
    static const char* names[] = {"INVALID", "FULL", "BFS", "MINIMUM", "NONE"};
    std::printf("pumipic buffer method %s\n", names[bufferMethod + 1]);
    std::printf("pumipic safe method %s\n", names[safeMethod + 1]);
  }
  static Method getMethod(std::string s) {   // case-insensitive, pumipic_input.cpp:139-151
    for (char& c : s) c = (char)std::toupper((unsigned char)c);
    return s == "FULL" ? FULL : s == "BFS" ? BFS : s == "MINIMUM" ? MINIMUM : s == "NONE" ? NONE : INVALID;
  }
  Ownership getRule() const { return ownership_rule; }
  const std::vector<lid_t>& getPartition() const { return partition; }
  int bridge_dim;        // entity dimension the BFS goes through (defaults to 0)
  int bufferBFSLayers;   // Method BFS: layers of the buffer (defaults to 3)
  int safeBFSLayers;     // Method BFS: layers of the safe zone (defaults to 1)
  friend class Mesh;

 private:
  void init(Method b, Method s) {            // pumipic_input.cpp:94-110
    bufferMethod = b == NONE ? MINIMUM : b;
    safeMethod = s;
    bridge_dim = 0;
    bufferBFSLayers = bufferMethod == MINIMUM ? 0 : 3;
    safeBFSLayers = safeMethod == MINIMUM ? 0 : 1;
  }
  // owner of every element: the partition itself, or through the elements' class ids
  // (setOwnerByClassification, part_construct.cpp:278-301)
  std::vector<lid_t> elementOwners() const {
    if (ownership_rule == PARTITION || resolved) return partition;
    const int dim = pp_host_mesh_dim(m), ne = pp_host_mesh_nents(m, dim);
    pp_host_tag t;
    pp_check(pp_host_mesh_find_tag(m, dim, "class_id", &t), "Input: class_id tag");
    const int32_t* cls = static_cast<const int32_t*>(t.data);
    std::vector<lid_t> own((size_t)ne);
    for (int e = 0; e < ne; ++e) {
      if (cls[e] < 0 || cls[e] >= (int)partition.size()) throw std::runtime_error("Input: class id outside the partition vector");
      own[(size_t)e] = partition[(size_t)cls[e]];
    }
    return own;
  }
  const pp_host_mesh* m;
  Ownership ownership_rule;
  std::vector<lid_t> partition;
  bool resolved = false;
  Method bufferMethod, safeMethod;
  pp_comm* comm;
};

// ---------------------------------------------------------------- Mesh (PICpart handle)
class ParticleBalancer;
class Mesh {
 public:
  enum Op { SUM_OP, MAX_OP, MIN_OP, BCAST_OP };   // pumipic_mesh.hpp:57-62
  // host arrays with Omega_h's entity numbering (coords, ask_elem_verts, ask_down(dim,dim-1),
  // ask_verts_of(dim-1), class_id); derived search data is built once on the device
  Mesh(int dim, const std::vector<double>& coords, const std::vector<int>& elem2verts,
       const std::vector<int>& elem2sides, const std::vector<int>& side2verts,
       const std::vector<int>& elem_class = std::vector<int>())
      : dim_(dim), comm_(nullptr) {
    pp_mesh_desc d;
    d.dim = dim;
    d.nverts = (int32_t)(coords.size() / dim);
    d.nelems = (int32_t)(elem2verts.size() / (dim + 1));
    d.nsides = (int32_t)(side2verts.size() / dim);
    d.coords = coords.data(); d.elem2verts = elem2verts.data(); d.elem2sides = elem2sides.data();
    d.side2verts = side2verts.data(); d.elem_class = elem_class.empty() ? nullptr : elem_class.data();
    d.memspace = PP_HOST;
    pp_check(pp_mesh_create(&d, nullptr, &h_), "Mesh");
    nents_[0] = d.nverts; nents_[dim - 1] = d.nsides; nents_[dim] = d.nelems;
  }
  // pumipic::Mesh(Input&) (pumipic_mesh.cpp / pumipic_part_construct.cpp:116-262): the PICpart of
  // this rank as pp_host_picpart_build / pp_host_picpart_read made it.  The record stays owned by
  // the caller and must outlive the Mesh.  With a communicator of more than one rank the comm-array
  // plans of setupComm (pumipic_comm.cpp:12-184) are built per entity dimension on first use.
  Mesh(const pp_host_picpart* record, pp_comm* comm) : dim_(0), comm_(nullptr) { adopt(record, comm, false); }
  // an empty Mesh for pumipic::read (pumipic_mesh.hpp:150-151) to fill
  Mesh() : dim_(0), comm_(nullptr) {}
  // Mesh(Input&) (part_construct.cpp:73-114): builds this rank's PICpart record and owns it
  explicit Mesh(Input& in) : dim_(0), comm_(nullptr) {
    build(in.m, in.elementOwners(), (int)in.bufferMethod, (int)in.safeMethod, in.bufferBFSLayers, in.safeBFSLayers,
          in.bridge_dim, in.comm);
  }
  // Mesh(o::Mesh&, o::LOs owners) (:43-53): every PICpart is the full mesh and all of it is safe
  Mesh(const pp_host_mesh* full, const std::vector<lid_t>& owners, pp_comm* comm) : dim_(0), comm_(nullptr) {
    build(full, owners, (int)Input::FULL, (int)Input::FULL, 3, 1, 0, comm);
  }
  // Mesh(o::Mesh&, owners, ghost_layers, safe_layers) (:55-71): vertex-bridged BFS buffers and safe zone
  Mesh(const pp_host_mesh* full, const std::vector<lid_t>& owners, int ghost_layers, int safe_layers, pp_comm* comm)
      : dim_(0), comm_(nullptr) {
    if (ghost_layers < safe_layers) throw std::runtime_error("Ghost layers must be >= safe layers");
    build(full, owners, (int)Input::BFS, (int)Input::BFS, ghost_layers, safe_layers, 0, comm);
  }
  // make this Mesh the PICpart `record`; with own = true the record is destroyed with the Mesh
  void adopt(const pp_host_picpart* record, pp_comm* comm, bool own) {
    if (h_ || record_) throw std::runtime_error("Mesh: already holds a mesh");
    const pp_host_mesh* m = pp_host_picpart_mesh(record);
    if (!m) throw std::runtime_error("Mesh: the PICpart record has no mesh");
    record_ = record;
    owns_record_ = false;   // until everything below has worked: on failure the caller keeps the record
    comm_ = comm;
    dim_ = pp_host_mesh_dim(m);
    for (int d = 0; d <= dim_; ++d) nents_[d] = pp_host_mesh_nents(m, d);
    pp_mesh_desc d;
    d.dim = dim_; d.nverts = nents_[0]; d.nelems = nents_[dim_]; d.nsides = nents_[dim_ - 1];
    d.coords = pp_host_mesh_coords(m);
    d.elem2verts = pp_host_mesh_ent2verts(m, dim_);
    d.elem2sides = pp_host_mesh_down(m, dim_);
    d.side2verts = pp_host_mesh_ent2verts(m, dim_ - 1);
    d.elem_class = hostTag<int32_t>(dim_, "class_id", false);
    d.memspace = PP_HOST;
    try {
      pp_check(pp_mesh_create(&d, nullptr, &h_), "Mesh");
      pp_check(pp_mesh_set_picpart(h_, hostTag<int32_t>(dim_, "safe"), hostTag<int32_t>(dim_, "ownership"),
                                   pp_host_picpart_rank(record), PP_HOST, nullptr), "Mesh: PICpart tags");
    } catch (...) {
      if (h_) pp_mesh_destroy(h_);
      h_ = nullptr; record_ = nullptr; comm_ = nullptr;
      throw;
    }
    owns_record_ = own;
  }
  ~Mesh() {
    for (int d = 0; d < 4; ++d) if (plan_[d]) pp_comm_plan_destroy(plan_[d]);
    if (h_) pp_mesh_destroy(h_);
    if (owns_record_ && record_) pp_host_picpart_destroy(const_cast<pp_host_picpart*>(record_));
  }
  Mesh(const Mesh&) = delete;
  Mesh& operator=(const Mesh&) = delete;
  int dim() const { return dim_; }
  lid_t nelems() const { return nents_[dim_]; }
  lid_t nents(int d) const { return nents_[d]; }
  pp_mesh* handle() const { return h_; }
  pp_mesh* operator->() const { return h_; }
  bool isFullMesh() const { return record_ ? pp_host_picpart_is_full_mesh(record_) != 0 : true; }
  // ---- the PICpart record (pumipic_mesh.hpp:33-52); only for a Mesh made from one
  const pp_host_picpart* record() const { return record_; }
  // Mesh::mesh(): the PICpart's own mesh with its tags (host side); null for a Mesh made from arrays
  const pp_host_mesh* mesh() const { return record_ ? pp_host_picpart_mesh(record_) : nullptr; }
  int numBuffers(int edim) const { return dimInfo(edim).num_cores + 1; }
  std::vector<lid_t> bufferedRanks(int edim) const {
    const pp_host_picpart_dim i = dimInfo(edim);
    return std::vector<lid_t>(i.buffered_parts, i.buffered_parts + (i.num_cores > 0 ? i.num_cores : 0));
  }
  View<gid_t> globalIds(int edim) const {
    const int64_t* g = hostTag<int64_t>(edim, "gids");
    return View<gid_t>(std::vector<gid_t>(g, g + nents_[edim]));
  }
  View<lid_t> safeTag() const { return tagView(dim_, "safe"); }
  View<lid_t> entOwners(int edim) const { return tagView(edim, "ownership"); }
  View<lid_t> rankLocalIndex(int edim) const { return tagView(edim, "rank_lids"); }
  View<lid_t> nentsOffsets(int edim) const {
    const pp_host_picpart_dim i = dimInfo(edim);
    return View<lid_t>(std::vector<lid_t>(i.offset_ents_per_rank,
                                          i.offset_ents_per_rank + pp_host_picpart_nranks(record_) + 1));
  }
  View<lid_t> commArrayIndex(int edim) const {
    const pp_host_picpart_dim i = dimInfo(edim);
    return View<lid_t>(std::vector<lid_t>(i.ent_to_comm_arr_index, i.ent_to_comm_arr_index + i.nents));
  }
  // PICpart tags (Mesh::safeTag(), Mesh::entOwners(dim)) and the communicator
  void setPICpart(const std::vector<int>& safe, const std::vector<int>& owners, int self_rank, pp_comm* comm) {
    pp_check(pp_mesh_set_picpart(h_, safe.data(), owners.data(), self_rank, PP_HOST, nullptr), "setPICpart");
    comm_ = comm;
  }
  pp_comm* comm() const { return comm_; }
  // Mesh::ptclBalancer() (pumipic_mesh.hpp:76).  The reference builds it lazily by exchanging safe
  // flags over MPI; here it is made from the PICpart's sbar table (ParticleBalancer below) and
  // attached by the caller, who keeps ownership.
  ParticleBalancer* ptclBalancer() const { return balancer_; }
  void setPtclBalancer(ParticleBalancer* b) { balancer_ = b; }
  template <class T> View<T> createCommArray(int edim, int nvals, T init) {   // pumipic_comm.cpp:187-192
    return View<T>((size_t)nents_[edim] * nvals, init);
  }
  // Mesh::reduceCommArray (pumipic_comm.cpp:223-440).  Arrays are indexed by the PICpart's own entity
  // numbering (the reference permutes them into its "comm array" order for the MPI staging,
  // commArrayIndex; that order is internal to it).  Full-mesh PICparts: one all-reduce, or the
  // owner's value; partially buffered ones: owner fan-in / fan-out over the plan of this dimension.
  template <class T>
  void reduceCommArray(int edim, Op op, View<T> array, const int* ent_owner_dev = nullptr) {
    if (!comm_ || pp_comm_size(comm_) == 1) return;
    const int nvals = (int)(array.size() / nents_[edim]);
    const int32_t dt = std::is_same<T, double>::value ? PP_FLOAT64 : std::is_same<T, float>::value ? PP_FLOAT32
                     : sizeof(T) == 8 ? PP_INT64 : PP_INT32;
    if (!record_ || isFullMesh()) {
      View<lid_t> owners;
      if (op == BCAST_OP && !ent_owner_dev && record_) { owners = entOwners(edim); ent_owner_dev = owners.data(); }
      pp_check(pp_comm_array_reduce(comm_, array.data(), nents_[edim], nvals, dt, (int32_t)op,
                                    ent_owner_dev, nullptr), "reduceCommArray");
      cuda_check(cudaStreamSynchronize(nullptr), "reduceCommArray");
      return;
    }
    if (!plan_[edim])
      pp_check(pp_comm_plan_create(comm_, nents_[edim], hostTag<int64_t>(edim, "gids"),
                                   hostTag<int32_t>(edim, "ownership"), PP_HOST, nullptr, &plan_[edim]),
               "reduceCommArray: setupComm");
    pp_check(pp_comm_plan_reduce(plan_[edim], array.data(), nvals, dt, (int32_t)op, nullptr), "reduceCommArray");
  }

 private:
  void build(const pp_host_mesh* full, const std::vector<lid_t>& owners, int bm, int sm, int bl, int sl, int bridge,
             pp_comm* comm) {
    const int nranks = comm ? pp_comm_size(comm) : 1, rank = comm ? pp_comm_rank(comm) : 0;
    if ((int)owners.size() != pp_host_mesh_nents(full, pp_host_mesh_dim(full)))
      throw std::runtime_error("Mesh: one owner per element is required");
    pp_host_picpart* rec = nullptr;
    pp_check(pp_host_picpart_build_bridged(full, owners.data(), nranks, rank, bm, sm, bl, sl, bridge, &rec), "Mesh");
    try {
      adopt(rec, comm, true);
    } catch (...) {
      pp_host_picpart_destroy(rec);
      throw;
    }
  }
  template <class T>
  const T* hostTag(int edim, const char* name, bool required = true) const {
    if (!record_) throw std::runtime_error("Mesh: not built from a PICpart record");
    pp_host_tag t;
    if (pp_host_mesh_find_tag(pp_host_picpart_mesh(record_), edim, name, &t) != PP_OK) {
      if (!required) return nullptr;
      throw std::runtime_error(std::string("Mesh: the PICpart record has no tag ") + name);
    }
    return static_cast<const T*>(t.data);
  }
  View<lid_t> tagView(int edim, const char* name) const {
    const int32_t* p = hostTag<int32_t>(edim, name);
    return View<lid_t>(std::vector<lid_t>(p, p + nents_[edim]));
  }
  pp_host_picpart_dim dimInfo(int edim) const {
    if (!record_) throw std::runtime_error("Mesh: not built from a PICpart record");
    pp_host_picpart_dim i;
    pp_check(pp_host_picpart_get(record_, edim, &i), "Mesh: PICpart record");
    return i;
  }
  int dim_;
  lid_t nents_[4] = {0, 0, 0, 0};
  pp_mesh* h_ = nullptr;
  pp_comm* comm_;
  const pp_host_picpart* record_ = nullptr;
  bool owns_record_ = false;
  pp_comm_plan* plan_[4] = {nullptr, nullptr, nullptr, nullptr};
  ParticleBalancer* balancer_ = nullptr;
};

// pumipic::write(picparts, prefix) / pumipic::read(lib, comm, prefix, &mesh) (pumipic_mesh.hpp:147-151,
// src/pumipic_file.cpp:45-205): <prefix>_<nranks>.ppm/ holding one .osh and one .ppm per rank.  The
// communicator gives the rank and the rank count (null = one rank); `mesh` must be empty and owns the
// record it reads.
inline void write(Mesh& picparts, const char* prefix) {
  if (!picparts.record()) throw std::runtime_error("write: the Mesh was not built from a PICpart record");
  pp_check(pp_host_picpart_write(picparts.record(), prefix), "write");
}
inline void read(pp_comm* comm, const char* prefix, Mesh* mesh) {
  const int nranks = comm ? pp_comm_size(comm) : 1, rank = comm ? pp_comm_rank(comm) : 0;
  pp_host_picpart* rec = nullptr;
  pp_check(pp_host_picpart_read(prefix, nranks, rank, &rec), "read");
  try {
    mesh->adopt(rec, comm, true);
  } catch (...) {
    pp_host_picpart_destroy(rec);
    throw;
  }
}

// ---------------------------------------------------------------- search
namespace detail {
template <class PS, class Seg3>
inline bool run_search(Mesh& mesh, PS* ptcls, int variant, Seg3 x_orig, Seg3 x_tgt, View<lid_t>& elem_ids,
                       bool requireIntersection, View<lid_t>* inter_faces, View<fp_t>* inter_points,
                       int inter_dim, int looplimit) {
  const size_t cap = (size_t)ptcls->capacity();
  pp_search_args a;
  a.variant = variant;
  a.x_orig = x_orig.data(); a.x_tgt = x_tgt.data(); a.stride = x_tgt.stride();
  a.elem_ids_empty = elem_ids.size() == 0;
  if (elem_ids.size() == 0) elem_ids = View<lid_t>(cap, (lid_t)-1, "elem_ids");   // tpp:504-509 (-1 = "use the row element")
  a.elem_ids = elem_ids.data();
  a.require_intersection = requireIntersection;
  if (inter_faces && (requireIntersection || variant == PP_SEARCH_3D_LEGACY || variant == PP_SEARCH_3D)) {
    if (inter_faces->size() < cap) *inter_faces = View<lid_t>(cap, (lid_t)-1, "inter_faces");   // tpp:538-539
    if (inter_points->size() < cap * inter_dim) *inter_points = View<fp_t>(cap * inter_dim, 0.0, "inter_points");
  }
  a.inter_faces = inter_faces ? inter_faces->data() : nullptr;
  a.inter_points = inter_points ? inter_points->data() : nullptr;
  a.looplimit = looplimit;
  pp_search_stats st;
  pp_check(pp_search_mesh(mesh.handle(), ptcls->handle(), &a, &st, (pp_stream)ptcls->stream()), "search_mesh");
  return st.found != 0;
}
}  // namespace detail

// adjacency.hpp:37-45 / adjacency.tpp:642
template <class ParticleType, typename Segment3d, typename SegmentInt>
bool search_mesh(Mesh& mesh, ParticleStructure<ParticleType>* ptcls, Segment3d x_ps_orig, Segment3d x_ps_tgt,
                 SegmentInt /*pids*/, View<lid_t>& elem_ids, bool requireIntersection, View<lid_t>& inter_faces,
                 View<fp_t>& inter_points, int looplimit = 0, int /*debug*/ = 0) {
  return detail::run_search(mesh, ptcls, PP_SEARCH_NEW, x_ps_orig, x_ps_tgt, elem_ids, requireIntersection,
                            &inter_faces, &inter_points, mesh.dim(), looplimit);
}
// adjacency.hpp:559-562 (legacy 3D: line-triangle + dual graph)
template <class ParticleType, typename Segment3d, typename SegmentInt>
bool search_mesh(Mesh& mesh, ParticleStructure<ParticleType>* ptcls, Segment3d x_ps_d, Segment3d xtgt_ps_d,
                 SegmentInt /*pid_d*/, View<lid_t>& elem_ids, View<fp_t>& xpoints_d, View<lid_t>& xface_id,
                 int looplimit = 0, int /*debug*/ = 0) {
  return detail::run_search(mesh, ptcls, PP_SEARCH_3D_LEGACY, x_ps_d, xtgt_ps_d, elem_ids, false, &xface_id,
                            &xpoints_d, 3, looplimit);
}
// adjacency.tpp:614-640 RemoveParticleOnGeometricModelExit: the stock handler of
// trace_particle_through_mesh -- a particle whose exit side is on the model boundary is done, and
// either keeps its element and records the side (requireIntersection) or leaves the domain.
template <typename ParticleType, typename Segment3d>
struct RemoveParticleOnGeometricModelExit {
  RemoveParticleOnGeometricModelExit(Mesh&, bool requireIntersection) : requireIntersection_(requireIntersection) {}
  void operator()(Mesh& mesh, ParticleStructure<ParticleType>* ptcls, View<lid_t>& elem_ids,
                  View<lid_t>& inter_faces, View<lid_t>& lastExit, View<fp_t>& inter_points,
                  View<lid_t>& ptcl_done, Segment3d /*x_ps_orig*/, Segment3d /*x_ps_tgt*/) const {
    pp_search_args a = {};
    a.variant = PP_SEARCH_NEW;
    a.elem_ids = elem_ids.data();
    a.require_intersection = requireIntersection_;
    a.inter_faces = inter_faces.data();
    a.inter_points = inter_points.data();
    pp_check(pp_trace_check_model_intersection(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(),
                                               lastExit.data(), (pp_stream)ptcls->stream()),
             "check_model_intersection");
  }

 private:
  bool requireIntersection_;
};

// adjacency.tpp:460-612 trace_particle_through_mesh with a user handler: `func` is called on the
// host once per walk iteration, between find_exit_face and set_new_element, as
//   func(mesh, ptcls, elem_ids, inter_faces, lastExit, inter_points, ptcl_done, x_ps_orig, x_ps_tgt)
// and works on the device arrays (its own kernels / parallel_for).  One kernel per phase, like the
// reference; search_mesh (one fused kernel) is the path for the stock handler.
template <class ParticleType, typename Segment3d, typename SegmentInt, typename Func>
bool trace_particle_through_mesh(Mesh& mesh, ParticleStructure<ParticleType>* ptcls, Segment3d x_ps_orig,
                                 Segment3d x_ps_tgt, SegmentInt /*pids*/, View<lid_t>& elem_ids,
                                 bool requireIntersection, View<lid_t>& inter_faces, View<fp_t>& inter_points,
                                 int looplimit, bool /*debug*/, Func& func) {
  const size_t cap = (size_t)ptcls->capacity();
  const int dim = mesh.dim();
  pp_search_args a = {};
  a.variant = PP_SEARCH_NEW;
  a.x_orig = x_ps_orig.data(); a.x_tgt = x_ps_tgt.data(); a.stride = x_ps_tgt.stride();
  a.elem_ids_empty = elem_ids.size() == 0;
  if (elem_ids.size() == 0) elem_ids = View<lid_t>(cap, "elem_ids");                           // :504-509
  a.elem_ids = elem_ids.data();
  a.require_intersection = requireIntersection;
  if (requireIntersection) {                                                                  // :535-549
    if (inter_faces.size() < cap) inter_faces = View<lid_t>(cap, (lid_t)-1, "inter_faces");
    if (inter_points.size() < cap * dim) inter_points = View<fp_t>(cap * dim, 0.0, "inter_points");
  }
  a.inter_faces = inter_faces.data();
  a.inter_points = inter_points.data();
  a.looplimit = looplimit;
  View<lid_t> ptcl_done(cap > 0 ? cap : 1, "ptcl_done"), lastExit(cap > 0 ? cap : 1, "lastExit");
  pp_stream st = (pp_stream)ptcls->stream();
  int32_t n = 0;
  pp_check(pp_trace_begin(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(), lastExit.data(), &n, st),
           "trace_particle_through_mesh: begin");
  if (n) std::fprintf(stderr, "[WARNING] %d particles are not in their parent element and were deleted\n", n);
  bool found = false;
  int loops = 0;
  while (!found) {
    pp_check(pp_trace_find_exit_face(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(), lastExit.data(), st),
             "find_exit_face");
    func(mesh, ptcls, elem_ids, inter_faces, lastExit, inter_points, ptcl_done, x_ps_orig, x_ps_tgt);
    pp_check(pp_trace_set_new_element(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(), lastExit.data(), st),
             "set_new_element");
    pp_check(pp_trace_pending(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(), lastExit.data(), 0, &n, st),
             "trace_particle_through_mesh: done check");
    found = n == 0;
    ++loops;
    if (loops > 1000000)   // a handler that never finishes its particles must not hang the caller
      throw std::runtime_error("trace_particle_through_mesh: no progress after 1000000 iterations");
    if (looplimit && loops >= looplimit) {                                                    // :584-606
      pp_check(pp_trace_pending(mesh.handle(), ptcls->handle(), &a, ptcl_done.data(), lastExit.data(), 1, &n, st),
               "trace_particle_through_mesh: loop limit");
      std::fprintf(stderr, "[ERROR] loop limit %d exceeded. %d particles were not found. Deleting them...\n",
                   looplimit, n);
      break;
    }
  }
  return found;
}

// adjacency.hpp:316-324 search_mesh_3d (GITRm's search: barycentric_coords_tet with tol 1e-20)
template <class ParticleStruct, typename CurrentCoordView, typename TargetCoordView, typename SegmentInt>
bool search_mesh_3d(Mesh& mesh, ParticleStruct* ptcls, CurrentCoordView x_ps_d, TargetCoordView xtgt_ps_d,
                    SegmentInt /*pid_d*/, View<lid_t>& elem_ids, View<fp_t>& xpoints_d, View<lid_t>& xface_d,
                    int looplimit = 0, int /*debug*/ = 0) {
  return detail::run_search(mesh, ptcls, PP_SEARCH_3D, x_ps_d, xtgt_ps_d, elem_ids, false, &xface_d,
                            &xpoints_d, 3, looplimit);
}
// adjacency.hpp:1013-1020 (elem_ids by value, -1 = start in the row element)
template <class ParticleType, typename Segment3d, typename SegmentInt>
bool search_mesh_2d(Mesh& mesh, ParticleStructure<ParticleType>* ptcls, Segment3d x_ps_d, Segment3d xtgt_ps_d,
                    SegmentInt /*pid_d*/, View<lid_t> elem_ids, int looplimit = 0, bool /*debug*/ = false) {
  return detail::run_search(mesh, ptcls, PP_SEARCH_2D_LEGACY, x_ps_d, xtgt_ps_d, elem_ids, false, nullptr,
                            nullptr, 2, looplimit);
}

// ---------------------------------------------------------------- migration ops (ptcl_ops.hpp)
template <class PS>
void setUnsafeProcs(Mesh& mesh, PS* ptcls, View<lid_t> elems, View<lid_t>& new_elems, View<lid_t>& new_procs) {
  const size_t cap = (size_t)ptcls->capacity();
  if (new_elems.size() < cap) new_elems = View<lid_t>(cap);
  if (new_procs.size() < cap) new_procs = View<lid_t>(cap);
  pp_check(pp_set_unsafe_procs(mesh.handle(), ptcls->handle(), elems.data(), new_elems.data(), new_procs.data(),
                               (pp_stream)ptcls->stream()), "setUnsafeProcs");
}
template <class PS>
void migrate_ptcls(Mesh& mesh, PS* ptcls, View<lid_t> new_elems) {
  View<lid_t> ptcl_elems, ptcl_procs;
  setUnsafeProcs(mesh, ptcls, new_elems, ptcl_elems, ptcl_procs);
  ptcls->migrate(ptcl_elems, ptcl_procs, Distributor(mesh.comm()));
}
// ---------------------------------------------------------------- particle load balancing (pumipic_lb.hpp)
// pumipic::ParticleBalancer (src/pumipic_lb.hpp:32-115).  Built from what the host PICpart record
// holds (pp_host_picpart_sbars, the "sbar_id" and "ownership" element tags) instead of the MPI
// exchange of the reference's constructor (pumipic_lb.cpp:23-82).
class ParticleBalancer {
 public:
  ParticleBalancer(Mesh& picparts, const std::vector<int>& sbar_ids, const std::vector<int>& parts_off,
                   const std::vector<int>& parts, const std::vector<int>& elem_sbar,
                   const std::vector<int>& elem_owner, int self_rank = 0) {
    pp_comm* c = picparts.comm();
    const int nranks = c ? pp_comm_size(c) : 1;
    const int rank = c ? pp_comm_rank(c) : self_rank;
    pp_check(pp_balancer_create(nranks, rank, (int32_t)sbar_ids.size(), sbar_ids.data(), parts_off.data(),
                                parts.data(), (int32_t)elem_sbar.size(), elem_sbar.data(), elem_owner.data(),
                                PP_HOST, c, nullptr, &h_), "ParticleBalancer");
    sbar_ids_ = View<lid_t>(elem_sbar);
  }
  // ParticleBalancer(Mesh& picparts) (pumipic_lb.cpp:23): everything comes from the Mesh's record
  explicit ParticleBalancer(Mesh& picparts) {
    const pp_host_picpart* rec = picparts.record();
    if (!rec) throw std::runtime_error("ParticleBalancer: the Mesh was not built from a PICpart record");
    int32_t n = 0, mx = 0;
    const int32_t *ids = nullptr, *off = nullptr, *parts = nullptr;
    pp_check(pp_host_picpart_sbars(rec, &n, &ids, &off, &parts, &mx), "ParticleBalancer: sbars");
    pp_host_tag sb, ow;
    const pp_host_mesh* m = pp_host_picpart_mesh(rec);
    pp_check(pp_host_mesh_find_tag(m, picparts.dim(), "sbar_id", &sb), "ParticleBalancer: sbar_id tag");
    pp_check(pp_host_mesh_find_tag(m, picparts.dim(), "ownership", &ow), "ParticleBalancer: ownership tag");
    pp_check(pp_balancer_create(pp_host_picpart_nranks(rec), pp_host_picpart_rank(rec), n, ids, off, parts,
                                picparts.nelems(), static_cast<const int32_t*>(sb.data),
                                static_cast<const int32_t*>(ow.data), PP_HOST, picparts.comm(), nullptr, &h_),
             "ParticleBalancer");
    const int32_t* p = static_cast<const int32_t*>(sb.data);
    sbar_ids_ = View<lid_t>(std::vector<lid_t>(p, p + picparts.nelems()));
  }
  ~ParticleBalancer() { if (h_) pp_balancer_destroy(h_); }
  ParticleBalancer(const ParticleBalancer&) = delete;
  ParticleBalancer& operator=(const ParticleBalancer&) = delete;
  pp_balancer* handle() const { return h_; }
  View<lid_t> getSbarIDs(Mesh&) const { return sbar_ids_; }                       // pumipic_lb.cpp:464-466
  // pumipic_lb.hpp:355-366; new_procs is changed in place
  template <class PS>
  void repartition(Mesh& picparts, PS* ps, double tol, View<lid_t> new_elems, View<lid_t> new_procs,
                   double step_factor = 0.3) {
    pp_check(pp_balancer_repartition(h_, picparts.comm(), ps->handle(), tol, new_elems.data(), new_procs.data(),
                                     step_factor, (pp_stream)ps->stream()), "ParticleBalancer::repartition");
  }
  // pumipic_lb.hpp:368-381: the new process of every particle, particles of element e at
  // [scan(e), scan(e) + ptcls_per_elem[e]); selection_iterations is accepted for source
  // compatibility (one pass assigns every planned send here)
  View<lid_t> partition(Mesh& picparts, View<lid_t> ptcls_per_elem, double tol, double step_factor = 0.3,
                        int /*selection_iterations*/ = 5) {
    const std::vector<lid_t> h = ptcls_per_elem.toHost();
    long np = 0;
    for (lid_t c : h) np += c;
    View<lid_t> new_procs((size_t)np, 0);
    pp_check(pp_balancer_partition(h_, picparts.comm(), ptcls_per_elem.data(), np, tol, step_factor,
                                   new_procs.data(), nullptr), "ParticleBalancer::partition");
    cuda_check(cudaStreamSynchronize(nullptr), "ParticleBalancer::partition");
    return new_procs;
  }
  // the steps of repartition, callable on their own (pumipic_lb.hpp:70-90)
  template <class PS>
  void addWeights(Mesh&, PS* ps, View<lid_t> new_elems, View<lid_t> new_procs) {
    pp_check(pp_balancer_add_weights_ps(h_, ps->handle(), new_elems.data(), new_procs.data(),
                                        (pp_stream)ps->stream()), "ParticleBalancer::addWeights");
  }
  void addWeights(Mesh&, View<lid_t> ptcls_per_elem) {
    pp_check(pp_balancer_add_weights_array(h_, ptcls_per_elem.data(), nullptr), "ParticleBalancer::addWeights");
  }
  // ParticlePlan stays inside the handle: balance() then selectParticles()
  void balance(Mesh& picparts, double tol, double step_factor = 0.3) {
    pp_check(pp_balancer_balance(h_, picparts.comm(), tol, step_factor, nullptr), "ParticleBalancer::balance");
  }
  template <class PS>
  void selectParticles(Mesh&, PS* ps, View<lid_t> new_elems, View<lid_t> new_parts) {
    pp_check(pp_balancer_select_ps(h_, ps->handle(), new_elems.data(), new_parts.data(),
                                   (pp_stream)ps->stream()), "ParticleBalancer::selectParticles");
  }

 private:
  pp_balancer* h_ = nullptr;
  View<lid_t> sbar_ids_;
};

// ptcl_ops.hpp:55-71: setUnsafeProcs -> ParticleBalancer::repartition -> migrate.  On one rank the
// balancer is a no-op (pumipic_lb.hpp:360-361); on several a balancer must have been attached.
template <class PS>
void migrate_lb_ptcls(Mesh& mesh, PS* ptcls, View<lid_t> new_elems, float tol, float step_factor = 0.5) {
  View<lid_t> ptcl_elems, ptcl_procs;
  setUnsafeProcs(mesh, ptcls, new_elems, ptcl_elems, ptcl_procs);
  if (mesh.comm() && pp_comm_size(mesh.comm()) > 1) {
    if (!mesh.ptclBalancer())
      throw std::runtime_error("migrate_lb_ptcls: no ParticleBalancer attached (Mesh::setPtclBalancer)");
    mesh.ptclBalancer()->repartition(mesh, ptcls, tol, ptcl_elems, ptcl_procs, step_factor);
  }
  ptcls->migrate(ptcl_elems, ptcl_procs, Distributor(mesh.comm()));
}

// ---------------------------------------------------------------- ViewComm (support/ViewComm.h:51-291)
// PS_Comm_* on device views over the NCCL communicator of the C ABI; `PS_Comm` stands where MPI_Comm
// stood, PS_Request where MPI_Request stood.  NCCL has no tags: messages between a pair of ranks
// match in posting order (the reference's call sites post one message per tag and peer, in the
// same order on both sides).  Non-blocking sends / receives are deferred and issued together --
// inside one ncclGroup -- by the first PS_Comm_Wait / PS_Comm_Waitall that needs one of them, which
// is also where the reference finishes its deferred unpacks (ViewComm.h:153-154).
typedef pp_comm* PS_Comm;
enum PS_Op { PS_SUM = PP_SUM, PS_MAX = PP_MAX, PS_MIN = PP_MIN };
struct PS_Request {
  int kind = 0;            // 0 done / empty, 1 send, 2 recv
  void* buf = nullptr;
  int64_t count = 0;
  int32_t dtype = 0;
  int peer = -1;
  pp_comm* comm = nullptr;
};
namespace detail {
template <class T> struct comm_dtype;
template <> struct comm_dtype<int> { static constexpr int32_t value = PP_INT32; };
template <> struct comm_dtype<long> { static constexpr int32_t value = PP_INT64; };
template <> struct comm_dtype<long long> { static constexpr int32_t value = PP_INT64; };
template <> struct comm_dtype<float> { static constexpr int32_t value = PP_FLOAT32; };
template <> struct comm_dtype<double> { static constexpr int32_t value = PP_FLOAT64; };
inline std::vector<PS_Request*>& pending_requests() {
  static std::vector<PS_Request*> p;
  return p;
}
// issue every deferred message of `comm` in one group, then wait for the stream
inline int flush_requests(pp_comm* comm) {
  std::vector<PS_Request*>& all = pending_requests();
  std::vector<PS_Request*> mine, rest;
  for (PS_Request* r : all) (r->comm == comm ? mine : rest).push_back(r);
  all.swap(rest);
  if (mine.empty()) return 0;
  pp_status st = pp_comm_group_start();
  for (PS_Request* r : mine) {
    if (st == PP_OK)
      st = r->kind == 1 ? pp_comm_send(comm, r->buf, r->count, r->dtype, r->peer, nullptr)
                        : pp_comm_recv(comm, r->buf, r->count, r->dtype, r->peer, nullptr);
    r->kind = 0;
  }
  const pp_status en = pp_comm_group_end();
  if (st == PP_OK) st = en;
  if (st == PP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) st = PP_ERR_CUDA;
  return (int)st;
}
}  // namespace detail

template <class T>
int PS_Comm_Send(View<T> view, int offset, int size, int dest, int /*tag*/, PS_Comm comm) {
  pp_status st = pp_comm_send(comm, view.data() + offset, size, detail::comm_dtype<T>::value, dest, nullptr);
  if (st == PP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) st = PP_ERR_CUDA;
  return (int)st;
}
template <class T>
int PS_Comm_Recv(View<T> view, int offset, int size, int source, int /*tag*/, PS_Comm comm) {
  pp_status st = pp_comm_recv(comm, view.data() + offset, size, detail::comm_dtype<T>::value, source, nullptr);
  if (st == PP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) st = PP_ERR_CUDA;
  return (int)st;
}
template <class T>
int PS_Comm_Isend(View<T> view, int offset, int size, int dest, int /*tag*/, PS_Comm comm, PS_Request* req) {
  req->kind = 1; req->buf = view.data() + offset; req->count = size;
  req->dtype = detail::comm_dtype<T>::value; req->peer = dest; req->comm = comm;
  detail::pending_requests().push_back(req);
  return 0;
}
template <class T>
int PS_Comm_Irecv(View<T> view, int offset, int size, int source, int /*tag*/, PS_Comm comm, PS_Request* req) {
  req->kind = 2; req->buf = view.data() + offset; req->count = size;
  req->dtype = detail::comm_dtype<T>::value; req->peer = source; req->comm = comm;
  detail::pending_requests().push_back(req);
  return 0;
}
inline int PS_Comm_Wait(PS_Request* req, void* /*status*/ = nullptr) {
  return req->kind ? detail::flush_requests(req->comm) : 0;
}
inline int PS_Comm_Waitall(int num_requests, PS_Request* requests, void* /*statuses*/ = nullptr) {
  int rc = 0;
  for (int i = 0; i < num_requests; ++i)
    if (requests[i].kind) { const int r = detail::flush_requests(requests[i].comm); if (r) rc = r; }
  return rc;
}
template <class T>
int PS_Comm_Alltoall(View<T> send_view, int send_size, View<T> recv_view, int /*recv_size*/, PS_Comm comm) {
  pp_status st = pp_comm_alltoall(comm, send_view.data(), recv_view.data(), send_size,
                                  detail::comm_dtype<T>::value, nullptr);
  if (st == PP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) st = PP_ERR_CUDA;
  return (int)st;
}
template <class T>
int PS_Comm_Ialltoall(View<T> send_view, int send_size, View<T> recv_view, int recv_size, PS_Comm comm,
                      PS_Request* req) {
  req->kind = 0;   // completed at once: the collective is already stream-ordered
  return PS_Comm_Alltoall(send_view, send_size, recv_view, recv_size, comm);
}
template <class T>
int PS_Comm_Allreduce(View<T> send_view, View<T> recv_view, int count, PS_Op op, PS_Comm comm) {
  pp_status st = pp_comm_allreduce(comm, send_view.data(), recv_view.data(), count,
                                   detail::comm_dtype<T>::value, (int32_t)op, nullptr);
  if (st == PP_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) st = PP_ERR_CUDA;
  return (int)st;
}
// the result is defined on `root` only, as with MPI_Reduce; the other ranks reduce into scratch
template <class T>
int PS_Comm_Reduce(View<T> send_view, View<T> recv_view, int count, PS_Op op, int root, PS_Comm comm) {
  if (pp_comm_rank(comm) == root) return PS_Comm_Allreduce(send_view, recv_view, count, op, comm);
  View<T> scratch((size_t)count);
  return PS_Comm_Allreduce(send_view, scratch, count, op, comm);
}

// printPtclImb (src/pumipic_lb.hpp:379-398): max, min, average and imbalance of the particle counts,
// printed by rank 0 (integer average first, like the reference's `tot_p / comm_size`)
template <class PS>
void printPtclImb(PS* ptcls, PS_Comm comm = nullptr) {
  const int np = ptcls->nPtcls();
  int min_p = np, max_p = np, tot_p = np, comm_size = 1, comm_rank = 0;
  if (comm && pp_comm_size(comm) > 1) {
    comm_size = pp_comm_size(comm);
    comm_rank = pp_comm_rank(comm);
    View<int> in(std::vector<int>{np}), out((size_t)1);
    PS_Comm_Allreduce(in, out, 1, PS_MIN, comm); min_p = out.toHost()[0];
    PS_Comm_Allreduce(in, out, 1, PS_MAX, comm); max_p = out.toHost()[0];
    PS_Comm_Allreduce(in, out, 1, PS_SUM, comm); tot_p = out.toHost()[0];
  }
  if (comm_rank == 0) {
    const float avg = (float)(tot_p / comm_size);
    const float imb = max_p / avg;
    std::printf("Ptcl LB <max, min, avg, imb>: %d %d %.3f %.3f\n", max_p, min_p, avg, imb);
  }
}

// ---------------------------------------------------------------- gather (field -> particle)
// Whole-structure forms of the interpolation helpers GITRm's push calls per particle
// (src/pumipic_adjacency.hpp:772-809, src/pumipic_utils.hpp:298-321,377-420,439-456): one kernel
// over all masked particles, output component-major like a particle member.
// interpolate3dFieldTet after findBCCoordsInTet; returns the number of particles outside their element
template <class PS, class Seg3>
int interpolate3dFieldTet(Mesh& mesh, PS* ptcls, Seg3 x, View<lid_t> elem_ids, View<fp_t> field, int dof,
                          View<fp_t>& out) {
  const size_t need = (size_t)dof * x.stride();
  if (out.size() < need) out = View<fp_t>(need, 0.0, "gathered field");
  int32_t bad = 0;
  pp_check(pp_gather_tet_field(mesh.handle(), ptcls->handle(), x.data(), x.stride(), elem_ids.data(), field.data(),
                               dof, out.data(), &bad, (pp_stream)ptcls->stream()), "interpolate3dFieldTet");
  return bad;
}
// interp2dVector on a uniform (r|x, z) grid; out is [3][stride]
template <class PS, class Seg3>
void interp2dVector(PS* ptcls, Seg3 x, View<fp_t> data3, fp_t gridx0, fp_t gridz0, fp_t dx, fp_t dz, lid_t nx,
                    lid_t nz, View<fp_t>& out, bool cylSymm = false) {
  const size_t need = (size_t)3 * x.stride();
  if (out.size() < need) out = View<fp_t>(need, 0.0, "gathered field");
  pp_check(pp_gather_grid2d_vector(ptcls->handle(), x.data(), x.stride(), data3.data(), gridx0, gridz0, dx, dz, nx,
                                   nz, cylSymm, out.data(), (pp_stream)ptcls->stream()), "interp2dVector");
}
// interpolate2d_field: one component of an nComp-component table; out is [stride]
template <class PS, class Seg3>
void interpolate2d_field(PS* ptcls, Seg3 x, View<fp_t> data, fp_t gridx0, fp_t gridz0, fp_t dx, fp_t dz, lid_t nx,
                         lid_t nz, View<fp_t>& out, bool cylSymm = true, lid_t nComp = 1, lid_t comp = 0) {
  if (out.size() < (size_t)x.stride()) out = View<fp_t>((size_t)x.stride(), 0.0, "gathered field");
  pp_check(pp_gather_grid2d(ptcls->handle(), x.data(), x.stride(), data.data(), gridx0, gridz0, dx, dz, nx, nz,
                            cylSymm, nComp, comp, out.data(), (pp_stream)ptcls->stream()), "interpolate2d_field");
}
// interpolate3d_field on grid lines gridx / gridy / gridz; out is [stride]
template <class PS, class Seg3>
void interpolate3d_field(PS* ptcls, Seg3 x, View<fp_t> gridx, View<fp_t> gridy, View<fp_t> gridz, View<fp_t> data,
                         View<fp_t>& out) {
  if (out.size() < (size_t)x.stride()) out = View<fp_t>((size_t)x.stride(), 0.0, "gathered field");
  pp_check(pp_gather_grid3d(ptcls->handle(), x.data(), x.stride(), data.data(), gridx.data(), gridy.data(),
                            gridz.data(), (int32_t)gridx.size(), (int32_t)gridy.size(), (int32_t)gridz.size(),
                            out.data(), (pp_stream)ptcls->stream()), "interpolate3d_field");
}

}  // namespace pumipic

namespace ps = pumipic;
