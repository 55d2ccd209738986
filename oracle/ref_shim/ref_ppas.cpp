// ref_ppas.cpp -- TEST INFRASTRUCTURE ONLY.  The constant-vector push and updatePtclPositions of the
// reference's test/pseudoPushAndSearch.cpp (:87-118, :142-154), extracted into
// ref_ppas.inc (a build-time temporary) and compiled unmodified.
#include "xgcm_shim.hpp"

namespace o = Omega_h;
namespace p = pumipic;
namespace ps = particle_structs;
using particle_structs::MemberTypes;
using pumipic::fp_t;
using pumipic::Vector3d;

namespace pumipic {
typedef KView<fp_t> kkFpView;                                   // src/pumipic_kktypes.hpp
static inline void hostToDeviceFp(kkFpView d, fp_t* h) { for (int i = 0; i < d.size(); ++i) d(i) = h[i]; }
}  // namespace pumipic
static inline void printTiming(const char*, double) {}

#include "ref_ppas.inc"

extern "C" {
void ref_push_constant(int cap, const int* slot_elem, const unsigned char* mask, double* x, double* xtgt, long stride,
                       double distance, double dx, double dy, double dz) {
  PS ptcls; pumipic::MemberViews mv;
  std::vector<int> ids((size_t)cap, 0);
  mv.arrays = {x, xtgt, ids.data()};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
  push(&ptcls, cap, distance, dx, dy, dz);
}
void ref_update_positions(int cap, const int* slot_elem, const unsigned char* mask, double* x, double* xtgt, long stride) {
  PS ptcls; pumipic::MemberViews mv;
  std::vector<int> ids((size_t)cap, 0);
  mv.arrays = {x, xtgt, ids.data()};
  mv.n = stride;
  ptcls.cap = cap; ptcls.slot_elem = slot_elem; ptcls.mask = mask; ptcls.members = &mv;
  updatePtclPositions(&ptcls);
}
}
