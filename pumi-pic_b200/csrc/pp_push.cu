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
}  // namespace

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
