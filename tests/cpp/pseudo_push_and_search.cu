// Driver written against the C++ mirror of the PUMI-PIC API (pumi-pic_b200/cpp/pumipic_b200.hpp),
// following the time-step loop of the reference's test/pseudoPushAndSearch.cpp:513-542 and the
// per-particle direction push of test/test_adj.cpp:550-562:
//   ps::parallel_for(push lambda) -> search_mesh -> ps::parallel_for(updatePtclPositions)
//   -> migrate_lb_ptcls (rebuild)
// Input (written by tests/test_cpp_mirror_gpu.py): cube size, particles, their elements, positions
// and directions.  Output: (pid, element, position) of every surviving particle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pumipic_b200.hpp"

namespace p = pumipic;
typedef p::MemberTypes<double[3], double[3], int, double[3]> Particle;   // test_adj.cpp:23 + direction
typedef p::ParticleStructure<Particle> PS;

template <class T> static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s in.bin out.bin nsteps kind(0 scs,1 csr,2 cabm,3 dps)\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  const int nsteps = atoi(argv[3]), kind = atoi(argv[4]);
  int hdr[2];
  if (fread(hdr, sizeof(int), 2, f) != 2) return 2;
  const int N = hdr[0], np = hdr[1];
  double dist;
  if (fread(&dist, sizeof(double), 1, f) != 1) return 2;
  // mesh: Kuhn cube with Omega_h-style derived sides
  int nv, ne, ns;
  double* co; int *ev, *e2s, *s2v;
  p::pp_check(pp_host_kuhn_cube(N, 1.0, &nv, &co, &ne, &ev), "kuhn");
  p::pp_check(pp_host_derive_sides(3, ne, ev, &ns, &e2s, &s2v), "sides");
  p::Mesh mesh(3, std::vector<double>(co, co + 3 * nv), std::vector<int>(ev, ev + 4 * ne),
               std::vector<int>(e2s, e2s + 4 * ne), std::vector<int>(s2v, s2v + 3 * ns));
  pp_host_free(co); pp_host_free(ev); pp_host_free(e2s); pp_host_free(s2v);
  std::vector<int> ppe = rd<int>(f, ne), pel = rd<int>(f, np);
  std::vector<double> x = rd<double>(f, 3 * (size_t)np), d = rd<double>(f, 3 * (size_t)np);
  fclose(f);
  std::vector<int> pid(np);
  for (int i = 0; i < np; ++i) pid[i] = i;
  std::vector<double> zero(3 * (size_t)np, 0.0);

  // particle_info: one device array per member, [ncomp][np]
  p::MemberTypeViews info = p::createMemberViews<Particle>(np);
  cudaMemcpy(info[0], x.data(), x.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(info[1], zero.data(), zero.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(info[2], pid.data(), pid.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(info[3], d.data(), d.size() * 8, cudaMemcpyHostToDevice);
  PS::kkLidView ppe_d(ppe), pel_d(pel);
  PS::kkGidView gids;
  p::TeamPolicy policy(10000, 32);                     // pseudoPushAndSearch.cpp:470-475
  PS* ptcls = nullptr;
  if (kind == 0) ptcls = new p::SellCSigma<Particle>(policy, INT_MAX, 1024, ne, np, ppe_d, gids, pel_d, info);
  else if (kind == 1) ptcls = new p::CSR<Particle>(policy, ne, np, ppe_d, gids, pel_d, info);
  else if (kind == 2) ptcls = new p::CabM<Particle>(policy, ne, np, ppe_d, gids, pel_d, info);
  else ptcls = new p::DPS<Particle>(policy, ne, np, ppe_d, gids, pel_d, info);
  p::destroyViews<Particle>(info);

  for (int iter = 1; iter <= nsteps; ++iter) {
    if (ptcls->nPtcls() == 0) break;
    auto x_ps = ptcls->get<0>();
    auto xtgt_ps = ptcls->get<1>();
    auto pid_ps = ptcls->get<2>();
    auto dir_ps = ptcls->get<3>();
    const double distance = dist;
    auto push = PS_LAMBDA(const int&, const int& slot, const bool& mask) {
      if (mask)
        for (int k = 0; k < 3; ++k) xtgt_ps(slot, k) = x_ps(slot, k) + distance * dir_ps(slot, k);
    };
    ps::parallel_for(ptcls, push, "push");
    p::View<p::lid_t> elem_ids, inter_faces;           // passed empty: allocated and seeded by the search
    p::View<p::fp_t> inter_points;
    const bool found = p::search_mesh(mesh, ptcls, x_ps, xtgt_ps, pid_ps, elem_ids, false, inter_faces,
                                      inter_points, 100);
    if (!found) { fprintf(stderr, "search failed\n"); return 1; }
    auto update = PS_LAMBDA(const int&, const int& slot, const bool&) {   // pseudoPushAndSearch.cpp:142-154
      for (int k = 0; k < 3; ++k) { x_ps(slot, k) = xtgt_ps(slot, k); xtgt_ps(slot, k) = 0; }
    };
    ps::parallel_for(ptcls, update, "updatePtclPositions");
    p::migrate_lb_ptcls(mesh, ptcls, elem_ids, 1.05f);
    printf("iter %d particles %d\n", iter, ptcls->nPtcls());
  }

  // dump (pid, row element, position) of the survivors
  pp_ps_layout lay;
  p::pp_check(pp_ps_get_layout(ptcls->handle(), nullptr, &lay), "layout");
  cudaDeviceSynchronize();
  const int cap = lay.capacity;
  std::vector<int> se(cap), pids(cap);
  std::vector<unsigned> mb((cap + 31) / 32 + 1);
  std::vector<double> pos(3 * (size_t)cap);
  auto x_ps = ptcls->get<0>();
  auto pid_ps = ptcls->get<2>();
  if (cap) {
    cudaMemcpy(se.data(), lay.slot_elem, cap * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(mb.data(), lay.mask_bits, ((cap + 31) / 32) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(pids.data(), pid_ps.data(), cap * 4, cudaMemcpyDeviceToHost);
    for (int k = 0; k < 3; ++k)
      cudaMemcpy(pos.data() + (size_t)k * cap, x_ps.data() + (size_t)k * x_ps.stride(), cap * 8, cudaMemcpyDeviceToHost);
  }
  FILE* o = fopen(argv[2], "wb");
  int n = 0;
  for (int s = 0; s < cap; ++s) n += (mb[s >> 5] >> (s & 31)) & 1u;
  fwrite(&n, 4, 1, o);
  for (int s = 0; s < cap; ++s)
    if ((mb[s >> 5] >> (s & 31)) & 1u) {
      fwrite(&pids[s], 4, 1, o); fwrite(&se[s], 4, 1, o);
      const double q[3] = {pos[s], pos[(size_t)cap + s], pos[2 * (size_t)cap + s]};
      fwrite(q, 8, 3, o);
    }
  fclose(o);
  delete ptcls;
  return 0;
}
