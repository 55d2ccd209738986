// pp_gather.cu -- mesh / grid -> particle field interpolation (the "gather" on the other side of
// the push).  One thread per slot, component-major SoA, coalesced particle
// columns; the field tables are read through the read-only path (they are small and L2-resident).
//
// Replaces the device helpers GITRm's push calls inside its ps::parallel_for lambdas:
//   src/pumipic_adjacency.hpp:772-809  interpolateTetVtx / interpolate3dFieldTet / findBCCoordsInTet
//   src/pumipic_utils.hpp:245-321      interpolate2d_base / interpolate2d / interpolate2d_field
//   src/pumipic_utils.hpp:377-420      interpolate3d_field
//   src/pumipic_utils.hpp:439-456      interp2dVector
// Arithmetic follows the reference operation by operation (-fmad=false), so results are
// bit-identical to the CPU oracle except where cos / sin / atan2 enter (cylindrical rotation).
#include "pp_internal.cuh"

namespace {
constexpr int kBlock = 256;

__device__ __forceinline__ bool slot_mask(const PsView& v, int s) {
  return (__ldg(v.mask_bits + (s >> 5)) >> (s & 31)) & 1u;
}

// findBCCoordsInTet (adjacency.hpp:801-809) = find_barycentric_tet (:97-133) on the gathered
// vertices, then interpolateTetVtx (:772-790) per component: bcc[fi] weighs the vertex opposite
// to face fi (simplex_opposite_template(3,2,fi) = 3,2,0,1).
__global__ void k_gather_tet_field(PsView v, const double* __restrict__ x, long stride,
                                   const int* __restrict__ elem_ids, const int* __restrict__ ev,
                                   const double* __restrict__ coords, const double* __restrict__ field,
                                   int dof, double* __restrict__ out, int* __restrict__ bad) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
  const int e = elem_ids[s];
  if (e < 0) return;
  const int4 tv = __ldg(reinterpret_cast<const int4*>(ev) + e);
  const int vid[4] = {tv.x, tv.y, tv.z, tv.w};
  d3 M[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    M[i] = {__ldg(coords + 3 * (long)vid[i]), __ldg(coords + 3 * (long)vid[i] + 1),
            __ldg(coords + 3 * (long)vid[i] + 2)};
  const d3 p = {x[s], x[stride + s], x[2 * stride + s]};
  const d3 n0 = cross3(M[1] - M[0], M[2] - M[0]);
  const d3 n1 = cross3(M[3] - M[0], M[1] - M[0]);
  const d3 n2 = cross3(M[3] - M[1], M[2] - M[1]);
  const d3 n3 = cross3(M[3] - M[2], M[0] - M[2]);
  const d3 p0 = p - M[0];
  double b[4];
  b[0] = dot3(p0, n0);
  b[1] = dot3(p0, n1);
  b[2] = dot3(p - M[1], n2);
  b[3] = dot3(p - M[2], n3);
  const double vol6 = dot3(M[3] - M[0], n0);
  bool ok = vol6 > 1.0e-20;
  if (ok) {
    const double inv = 1.0 / vol6;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      b[i] = inv * b[i];
      ok = ok && pp_gtez(b[i], 1e-10);
    }
  }
  if (!ok) { atomicAdd(bad, 1); return; }   // OMEGA_H_CHECK in the reference
  for (int c = 0; c < dof; ++c) {
    double val = 0;
    val = val + b[0] * __ldg(field + (long)vid[3] * dof + c);
    val = val + b[1] * __ldg(field + (long)vid[2] * dof + c);
    val = val + b[2] * __ldg(field + (long)vid[0] * dof + c);
    val = val + b[3] * __ldg(field + (long)vid[1] * dof + c);
    out[c * stride + s] = val;
  }
}

// pumipic_utils.hpp:245-248
__device__ __forceinline__ double interp_base(double d1, double d2, double g1, double g2, double v,
                                              double dv) {
  return (d1 * (g2 - v) + d2 * (v - g1)) / dv;
}

struct Grid2 {
  double x0, z0, dx, dz;
  int nx, nz;
};

// pumipic_utils.hpp:298-321 interpolate2d_field -> :260-296 interpolate2d (cylSymm already applied)
__device__ __forceinline__ double interp2d_field(const double* __restrict__ data, const Grid2& g,
                                                 double x, double z, int nComp, int comp) {
  if (g.nx <= 1 && g.nz <= 1) return __ldg(data + comp);
  int i = (int)floor((x - g.x0) / g.dx);
  int j = (int)floor((z - g.z0) / g.dz);
  if (i < 0) i = 0;
  if (j < 0) j = 0;
  const double gXi = g.x0 + i * g.dx, gXip1 = g.x0 + (i + 1) * g.dx;
  const double gZj = g.z0 + j * g.dz, gZjp1 = g.z0 + (j + 1) * g.dz;
  const int nx = g.nx, nz = g.nz;
  auto D = [&](long idx) { return __ldg(data + idx * nComp + comp); };
  if (i >= nx - 1 && j >= nz - 1) return D(nx - 1 + (long)(nz - 1) * nx);
  if (i >= nx - 1)
    return interp_base(D(nx - 1 + (long)j * nx), D(nx - 1 + (long)(j + 1) * nx), z - gZj, gZjp1 - z, z, g.dz);
  if (j >= nz - 1)
    return interp_base(D(i + (long)(nz - 1) * nx), D(i + (long)(nz - 1) * nx), x - gXi, gXip1 - x, x, g.dx);
  const double f1 = interp_base(D(i + (long)j * nx), D(i + 1 + (long)j * nx), gXi, gXip1, x, g.dx);
  const double f2 = interp_base(D(i + (long)(j + 1) * nx), D(i + 1 + (long)(j + 1) * nx), gXi, gXip1, x, g.dx);
  return interp_base(f1, f2, gZj, gZjp1, z, g.dz);
}

// ncomp_out == 1: interpolate2d_field of component `comp`; ncomp_out == 3: interp2dVector (:439-456)
__global__ void k_gather_grid2d(PsView v, const double* __restrict__ x, long stride,
                                const double* __restrict__ data, Grid2 g, int cyl, int nComp, int comp,
                                int vector3, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
  const double px = x[s], py = x[stride + s], pz = x[2 * stride + s];
  double r = px;
  if (cyl) r = sqrt(px * px + py * py);
  if (!vector3) {
    out[s] = interp2d_field(data, g, r, pz, nComp, comp);
    return;
  }
  double f0 = interp2d_field(data, g, r, pz, 3, 0);
  double f1 = interp2d_field(data, g, r, pz, 3, 1);
  const double f2 = interp2d_field(data, g, r, pz, 3, 2);
  if (cyl) {
    const double theta = atan2(py, px);
    const double c = cos(theta), sn = sin(theta);
    const double a0 = f0, a1 = f1;
    f0 = c * a0 - sn * a1;
    f1 = sn * a0 + c * a1;
  }
  out[s] = f0; out[stride + s] = f1; out[2 * stride + s] = f2;
}

// pumipic_utils.hpp:377-420 interpolate3d_field
__global__ void k_gather_grid3d(PsView v, const double* __restrict__ x, long stride,
                                const double* __restrict__ data, const double* __restrict__ gx,
                                const double* __restrict__ gy, const double* __restrict__ gz, int nx,
                                int ny, int nz, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.capacity || !slot_mask(v, s)) return;
  const double px = x[s], py = x[stride + s], pz = x[2 * stride + s];
  const double dx = __ldg(gx + 1) - __ldg(gx);
  const double dy = ny > 1 ? __ldg(gy + 1) - __ldg(gy) : 1.0;
  const double dz = nz > 1 ? __ldg(gz + 1) - __ldg(gz) : 1.0;
  int i = (int)floor((px - __ldg(gx)) / dx);
  int j = (int)floor((py - __ldg(gy)) / dy);
  int k = (int)floor((pz - __ldg(gz)) / dz);
  i = (i < 0) ? 0 : ((i >= nx - 1) ? (nx - 2) : i);
  j = (j < 0 || ny <= 1) ? 0 : ((j >= ny - 1) ? (ny - 2) : j);
  k = (k < 0 || nz <= 1) ? 0 : ((k >= nz - 1) ? (nz - 2) : k);
  const long nxy = (long)nx * ny;
  const double gxi = __ldg(gx + i), gxi1 = __ldg(gx + i + 1);
  auto row = [&](long idx) { return interp_base(__ldg(data + idx), __ldg(data + idx + 1), gxi, gxi1, px, dx); };
  const double fx_z0 = row(i + (long)j * nx + k * nxy);
  // the reference evaluates all four rows unconditionally; with ny <= 1 or nz <= 1 the extra rows
  // lie outside the table and their values are discarded, so they are skipped here
  double fxyz = fx_z0;
  if (nz > 1) {
    const double fx_z1 = row(i + (long)j * nx + (k + 1) * nxy);
    const double gzk = __ldg(gz + k), gzk1 = __ldg(gz + k + 1);
    const double fxz0 = interp_base(fx_z0, fx_z1, gzk, gzk1, pz, dz);
    fxyz = fxz0;
    if (ny > 1) {
      const double fxy_z0 = row(i + (long)(j + 1) * nx + k * nxy);
      const double fxy_z1 = row(i + (long)(j + 1) * nx + (k + 1) * nxy);
      const double fxz1 = interp_base(fxy_z0, fxy_z1, gzk, gzk1, pz, dz);
      fxyz = interp_base(fxz0, fxz1, __ldg(gy + j), __ldg(gy + j + 1), py, dy);
    }
  }
  out[s] = fxyz;
}
}  // namespace

extern "C" pp_status pp_gather_tet_field(pp_mesh* mesh, pp_ps* ps, const double* x, int64_t stride,
                                         const int32_t* elem_ids, const double* field, int32_t dof,
                                         double* out, int32_t* n_outside_host, pp_stream stream) {
  PP_REQUIRE(mesh && ps && x && elem_ids && field && out, "null argument");
  PP_REQUIRE(mesh->dim == 3, "interpolateTetVtx needs a 3D mesh");
  PP_REQUIRE(dof >= 1, "dof must be positive");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  cudaStream_t s = (cudaStream_t)stream;
  if (n_outside_host) *n_outside_host = 0;
  if (ps->capacity == 0) return PP_OK;
  int* bad;
  PP_TRY(pp_dev_alloc(&bad, 1, s));
  PP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
  k_gather_tet_field<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, s>>>(
      ps->view(), x, stride, elem_ids, mesh->elem2verts, mesh->coords, field, dof, out, bad);
  PP_KERNEL_CHECK();
  if (n_outside_host) {
    PP_CUDA(cudaMemcpyAsync(n_outside_host, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    PP_CUDA(cudaStreamSynchronize(s));
  }
  pp_dev_free(bad, s);
  return PP_OK;
}

extern "C" pp_status pp_gather_grid2d(pp_ps* ps, const double* x, int64_t stride, const double* data,
                                      double gridx0, double gridz0, double dx, double dz, int32_t nx,
                                      int32_t nz, int32_t cyl_symm, int32_t ncomp, int32_t comp,
                                      double* out, pp_stream stream) {
  PP_REQUIRE(ps && x && data && out, "null argument");
  PP_REQUIRE(nx >= 1 && nz >= 1 && ncomp >= 1 && comp >= 0 && comp < ncomp, "bad grid / component");
  PP_REQUIRE(dx > 0 && dz > 0, "dx and dz must be positive (OMEGA_H_CHECK in pumipic_utils.hpp:309)");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  const Grid2 g = {gridx0, gridz0, dx, dz, nx, nz};
  k_gather_grid2d<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, stride, data, g, cyl_symm ? 1 : 0, ncomp, comp, 0, out);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_gather_grid2d_vector(pp_ps* ps, const double* x, int64_t stride,
                                             const double* data3, double gridx0, double gridz0,
                                             double dx, double dz, int32_t nx, int32_t nz,
                                             int32_t cyl_symm, double* out, pp_stream stream) {
  PP_REQUIRE(ps && x && data3 && out, "null argument");
  PP_REQUIRE(nx >= 1 && nz >= 1, "bad grid");
  PP_REQUIRE(dx > 0 && dz > 0, "dx and dz must be positive (OMEGA_H_CHECK in pumipic_utils.hpp:309)");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  const Grid2 g = {gridx0, gridz0, dx, dz, nx, nz};
  k_gather_grid2d<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, stride, data3, g, cyl_symm ? 1 : 0, 3, 0, 1, out);
  PP_KERNEL_CHECK();
  return PP_OK;
}

extern "C" pp_status pp_gather_grid3d(pp_ps* ps, const double* x, int64_t stride, const double* data,
                                      const double* gridx, const double* gridy, const double* gridz,
                                      int32_t nx, int32_t ny, int32_t nz, double* out,
                                      pp_stream stream) {
  PP_REQUIRE(ps && x && data && gridx && gridy && gridz && out, "null argument");
  PP_REQUIRE(nx >= 2 && ny >= 1 && nz >= 1, "interpolate3d_field needs nx >= 2");
  PP_REQUIRE(stride >= ps->capacity, "stride smaller than capacity");
  if (ps->capacity == 0) return PP_OK;
  k_gather_grid3d<<<pp_div_up(ps->capacity, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
      ps->view(), x, stride, data, gridx, gridy, gridz, nx, ny, nz, out);
  PP_KERNEL_CHECK();
  return PP_OK;
}
