// pp_push.cu -- particle push kernels (one thread per slot, component-major SoA, coalesced).
//
// Replaces the driver push lambdas the reference runs through ps::parallel_for:
//   test/pseudoPushAndSearch.cpp:104-114 (constant vector), test/test_adj.cpp:550-562
//   (per-particle direction), test/pseudoPushAndSearch.cpp:142-154 (updatePtclPositions).
#include "pp_internal.cuh"

namespace {
constexpr int kBlock = 256;

__device__ __forceinline__ bool slot_mask(const PsView& v, int s) {
  return (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
}

__global__ void k_push_constant(PsView v, const double* __restrict__ x, double* __restrict__ xt,
                                long stride, double d0, double d1, double d2, double d3_) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
  // dir[i] = disp[0]*disp[i+1]; xtgt = x + dir + ptclUnique (a zero-filled array, :100)
  const double unique = 0.0;
  xt[s] = x[s] + d0 * d1 + unique;
  xt[stride + s] = x[stride + s] + d0 * d2 + unique;
  xt[2 * stride + s] = x[2 * stride + s] + d0 * d3_ + unique;
}

__global__ void k_push_direction(PsView v, double* __restrict__ tgt, const double* __restrict__ dir,
                                 long stride, double distance) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) tgt[i * stride + s] = tgt[i * stride + s] + distance * dir[i * stride + s];
}

__global__ void k_push_from(PsView v, const double* __restrict__ x, double* __restrict__ xt,
                            const double* __restrict__ dir, long stride, double distance) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) xt[i * stride + s] = x[i * stride + s] + distance * dir[i * stride + s];
}

__global__ void k_update_positions(int cap, double* __restrict__ x, double* __restrict__ xt, long stride) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;   // the reference ignores the mask here
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x[i * stride + s] = xt[i * stride + s];
    xt[i * stride + s] = 0;
  }
}

// src/pumipic_push.hpp:26-71 pushBoris, one thread per particle (the reference launches it with
// parallel_for(1, ...), :74, i.e. as a formula; here it runs over all n particles).
__global__ void k_push_boris(long n, long stride, double* __restrict__ pos, double* __restrict__ prev,
                             double* __restrict__ vel, const double* __restrict__ ef,
                             const double* __restrict__ bf, double qPrime, double dt) {
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const d3 v = {vel[p], vel[stride + p], vel[2 * stride + p]};
  const d3 E = {ef[p], ef[stride + p], ef[2 * stride + p]};
  const d3 B = {bf[p], bf[stride + p], bf[2 * stride + p]};
  const double bmag = norm3(B);
  const double coeff = 2.0 * qPrime / (1.0 + (qPrime * bmag) * (qPrime * bmag));
  const d3 qpE = {E.x * qPrime, E.y * qPrime, E.z * qPrime};
  const d3 vMinus = v - qpE;                                   // v_minus = v - q'E
  const d3 c1 = cross3(vMinus, B);
  const d3 vPrime = {vMinus.x + c1.x * qPrime, vMinus.y + c1.y * qPrime, vMinus.z + c1.z * qPrime};
  const d3 c2 = cross3(vPrime, B);
  d3 w = {vMinus.x + c2.x * coeff, vMinus.y + c2.y * coeff, vMinus.z + c2.z * coeff};
  w = {w.x + qpE.x, w.y + qpE.y, w.z + qpE.z};
  const d3 pre = {prev[p], prev[stride + p], prev[2 * stride + p]};
  prev[p] = pos[p]; prev[stride + p] = pos[stride + p]; prev[2 * stride + p] = pos[2 * stride + p];
  pos[p] = pre.x + w.x * dt; pos[stride + p] = pre.y + w.y * dt; pos[2 * stride + p] = pre.z + w.z * dt;
  vel[p] = w.x; vel[stride + p] = w.y; vel[2 * stride + p] = w.z;
}
}  // namespace

extern "C" pp_status pp_push_boris(int64_t n, int64_t stride, double* pos, double* pos_prev, double* vel,
                                   const double* efield, const double* bfield, double dt,
                                   pp_stream stream) {
  PP_REQUIRE(pos && pos_prev && vel && efield && bfield, "null argument");
  PP_REQUIRE(n >= 0 && stride >= n, "stride smaller than n");
  PP_REQUIRE(dt > 0, "dt must be positive (OMEGA_H_CHECK in pumipic_push.hpp:35)");
  if (n == 0) return PP_OK;
  const double charge = 1, amu = 10;   // pumipic_push.hpp:32-33
  const double qPrime = charge * 1.60217662e-19 / (amu * 1.6737236e-27) * dt * 0.5;
  k_push_boris<<<pp_div_up(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(n, stride, pos, pos_prev, vel,
                                                                           efield, bfield, qPrime, dt);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_push_constant(pp_ps* ps, const double* x, double* xtgt, int64_t stride,
                                      double distance, double dx, double dy, double dz,
                                      pp_stream stream) {
  PP_REQUIRE(ps && x && xtgt, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_push_constant<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, xtgt, stride, distance, dx, dy, dz);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_push_direction(pp_ps* ps, double* tgt, const double* dir, int64_t stride,
                                       double distance, pp_stream stream) {
  PP_REQUIRE(ps && tgt && dir, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_push_direction<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), tgt, dir, stride, distance);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_update_positions(pp_ps* ps, double* x, double* xtgt, int64_t stride,
                                         pp_stream stream) {
  PP_REQUIRE(ps && x && xtgt, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_update_positions<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->capacity, x, xtgt, stride);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_push_from(pp_ps* ps, const double* x, double* xtgt, const double* dir,
                                  int64_t stride, double distance, pp_stream stream) {
  PP_REQUIRE(ps && x && xtgt && dir, "null argument");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_push_from<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, xtgt, dir, stride, distance);
  PP_KERNEL_CHECK();
  return PP_OK;
}
