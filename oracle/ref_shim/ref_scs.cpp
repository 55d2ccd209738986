// ref_scs.cpp -- TEST INFRASTRUCTURE ONLY.  The geometry of the reference's Sell-C-sigma
// structure -- chooseChunkHeight, constructChunks (chunk widths with the three padding
// strategies) and constructOffsets (vertical slices, offsets, capacity) of
// particle_structs/src/scs/SCS_buildFns.h:4-153 -- compiled unmodified (extracted into ref_scs.inc,
// a build-time temporary) as members of a stand-in SellCSigma class that declares just what they
// touch.  The sigma sort in front of them (SCS_sort.h, thrust / Kokkos sort_by_key) is restated here:
// ascending by particle count inside windows of sigma rows; the order of equal counts, which the
// reference leaves to its sort backend, cannot influence widths, offsets or capacity.
#include <algorithm>
#include <numeric>

#include "omega_h_mesh_shim.hpp"

#define KOKKOS_LAMBDA [=]

namespace Kokkos {
struct ViewAllocateWithoutInitializing {
  explicit ViewAllocateWithoutInitializing(const char*) {}
};
template <class T, class Space> struct Max {
  T& ref;
  explicit Max(T& r) : ref(r) {}
};
struct HostSpace {};
template <class... A> struct RangePolicy {
  int begin, end;
  RangePolicy(int b, int e) : begin(b), end(e) {}
};
// one "team" per chunk; members run one after the other, team_reduce(Max) keeps the running maximum,
// so that the value the LAST member holds (and writes last) is the maximum over the team
struct TeamMember {
  int league, rank;
  int* running;
  int league_rank() const { return league; }
  int team_rank() const { return rank; }
  template <class T, class S> void team_reduce(Max<T, S> m) const {
    if (m.ref > *running) *running = m.ref;
    m.ref = *running;
  }
};
struct TeamPolicyStub {
  typedef TeamMember member_type;
  int league, team;
  TeamPolicyStub(int l, int t) : league(l), team(t) {}
};
template <class F> void parallel_for(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
template <class... A, class F> void parallel_for(RangePolicy<A...> r, F f) { for (int i = r.begin; i < r.end; ++i) f(i); }
template <class F> void parallel_for(const TeamPolicyStub& p, F f) {
  for (int l = 0; l < p.league; ++l) {
    int running = 0;
    for (int r = 0; r < p.team; ++r) f(TeamMember{l, r, &running});
  }
}
template <class F, class T> void parallel_reduce(const char*, int n, F f, T& out) {
  T v = T();
  for (int i = 0; i < n; ++i) f(i, v);
  out = v;
}
template <class T, class U> T atomic_fetch_add(T* p, U v) { T old = *p; *p += v; return old; }
}  // namespace Kokkos

namespace pumipic {
typedef long gid_t;
enum PaddingStrategy { PAD_EVENLY, PAD_PROPORTIONALLY, PAD_INVERSELY };   // scs_input.hpp
template <class T> struct ScsView : public KView<T> {                     // + the uninitialised constructor
  ScsView() {}
  ScsView(const std::string& s, int n) : KView<T>(s, n) {}
  ScsView(Kokkos::ViewAllocateWithoutInitializing, int n) : KView<T>("", n) {}
};
template <class V> int getLastValue(V v) { return v.size() ? v(v.size() - 1) : 0; }   // SupportKK.h:102-109
struct ExecSpace {};
template <class V, class S> void exclusive_scan(V in, V out, S) {         // SupportKK.h:14-16
  int acc = 0;
  for (int i = 0; i < in.size(); ++i) { const int x = in(i); out(i) = acc; acc += x; }
}
// the members constructChunks / constructOffsets / chooseChunkHeight read and write
template <class DataTypes, typename MemSpace>
class SellCSigma {
 public:
  typedef ScsView<lid_t> kkLidView;
  typedef ScsView<gid_t> kkGidView;
  typedef Kokkos::TeamPolicyStub PolicyType;
  typedef ExecSpace execution_space;
  lid_t num_elems = 0, C_ = 1, V_ = 1, num_empty_elements = 0;
  double shuffle_padding = 0;
  PaddingStrategy pad_strat = PAD_EVENLY;
  int chooseChunkHeight(int maxC, kkLidView ptcls_per_elem);
  void constructChunks(kkLidView ptcls, kkLidView index, lid_t& nchunks, kkLidView& chunk_widths,
                       kkLidView& row_element, kkLidView& element_row);
  void constructOffsets(lid_t nChunks, lid_t& nSlices, kkLidView chunk_widths, kkLidView& offs, kkLidView& s2c,
                        lid_t& cap);
};
#include "ref_scs.inc"
}  // namespace pumipic

// ppe[ne] -> the layout the reference builds for it.  Outputs sized by the caller: chunk_widths
// [ceil(ne / C)], offsets [nslices_max + 1], s2c [nslices_max]; returns the capacity.
extern "C" int ref_scs_layout(int ne, const int* ppe, int maxC, int sigma, int V, double shuffle_padding, int pad_strat,
                              int* C_out, int* nchunks_out, int* chunk_widths, int* nslices_out, int* offsets,
                              int* s2c, int max_slices, int* num_empty_out) {
  typedef pumipic::SellCSigma<pumipic::MemberTypes<int>, Kokkos::HostSpace> SCS;
  SCS scs;
  SCS::kkLidView ppe_v("ppe", ne);
  for (int i = 0; i < ne; ++i) ppe_v(i) = ppe[i];
  scs.num_elems = ne;
  scs.C_ = scs.chooseChunkHeight(maxC, ppe_v);
  scs.V_ = V;
  scs.shuffle_padding = shuffle_padding;
  scs.pad_strat = (pumipic::PaddingStrategy)pad_strat;
  // sigmaSort (SCS_sort.h:4-60): ascending inside windows of sigma rows, the last window takes the rest
  SCS::kkLidView ptcls("ptcls", ne), index("index", ne);
  std::vector<int> order((size_t)ne);
  std::iota(order.begin(), order.end(), 0);
  if (sigma > 1) {
    int i = 0;
    for (; (long)i < (long)ne - sigma; i += sigma)
      std::stable_sort(order.begin() + i, order.begin() + i + sigma, [&](int a, int b) { return ppe[a] < ppe[b]; });
    std::stable_sort(order.begin() + i, order.end(), [&](int a, int b) { return ppe[a] < ppe[b]; });
  }
  for (int i = 0; i < ne; ++i) { ptcls(i) = ppe[order[(size_t)i]]; index(i) = order[(size_t)i]; }
  pumipic::lid_t nchunks = 0, nslices = 0, cap = 0;
  SCS::kkLidView widths, row_element, element_row, offs, s2c_v;
  scs.constructChunks(ptcls, index, nchunks, widths, row_element, element_row);
  scs.constructOffsets(nchunks, nslices, widths, offs, s2c_v, cap);
  *C_out = scs.C_; *nchunks_out = nchunks; *nslices_out = nslices; *num_empty_out = scs.num_empty_elements;
  for (int i = 0; i < nchunks; ++i) chunk_widths[i] = widths(i);
  if (nslices <= max_slices) {
    for (int i = 0; i <= nslices; ++i) offsets[i] = offs(i);
    for (int i = 0; i < nslices; ++i) s2c[i] = s2c_v(i);
  }
  return cap;
}
