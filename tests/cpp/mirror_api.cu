// Self-checking driver for the parts of the C++ mirror (pumi-pic_b200/cpp/pumipic_b200.hpp) that the
// pseudoPushAndSearch driver does not touch, written like the reference's own tests:
//   getMemberView / createMemberViews   particle_structs/test (MemberTypeLibraries.h:33-41,90-105)
//   getPIDs                             particle_structs/test/test_structure.cpp:354-378 (testPIDs)
//   pumipic::Mesh(Input&) accessors     test/test_input_construct.cpp, test/test_comm_array.cpp
//   setUnsafeProcs                      src/pumipic_ptcl_ops.hpp:33-53
//   ParticleBalancer                    test/test_lb.cpp:78-130 (one process acting as rank 0 of 4)
//   PS_Comm_*                           support/ViewComm_test.cpp (one rank: self-consistent copies)
//   trace_particle_through_mesh         src/pumipic_adjacency.tpp:460-640 with a user handler
// Exit code 0 and "MIRROR_API_OK" on success.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "pumipic_b200.hpp"

namespace p = pumipic;
typedef p::MemberTypes<int, double[3], int> Particle;   // id, vector, int (test particle of particle_structs/test)
typedef p::ParticleStructure<Particle> PS;

static int fails = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { ++fails; fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

int main() {
  // ---------------------------------------------------------------- full mesh + 4 block owners
  const int N = 16, NR = 4;
  int nv, ne;
  double* co; int* ev;
  p::pp_check(pp_host_plate(N, 1.0, &nv, &co, &ne, &ev), "plate");
  pp_host_mesh* full = nullptr;
  p::pp_check(pp_host_mesh_from_elems(2, nv, co, ne, ev, &full), "from_elems");
  std::vector<int> owner(ne);
  for (int e = 0; e < ne; ++e) {
    double cx = 0, cy = 0;
    for (int k = 0; k < 3; ++k) { cx += co[2 * ev[3 * e + k]]; cy += co[2 * ev[3 * e + k] + 1]; }
    owner[e] = (cx / 3 >= 0.5 ? 1 : 0) + (cy / 3 >= 0.5 ? 2 : 0);
  }
  pp_host_picpart* rec = nullptr;   // Input(mesh, PARTITION, owner, BFS buffers, BFS safe zone of 2 layers)
  p::pp_check(pp_host_picpart_build(full, owner.data(), NR, 0, 1, 1, -1, 2, &rec), "picpart_build");

  // ---------------------------------------------------------------- Mesh from the record
  p::Mesh picparts(rec, nullptr);
  const int nel = picparts.nelems();
  CHECK(picparts.dim() == 2 && nel > 0 && nel <= ne && !picparts.isFullMesh());
  std::vector<int> safe = picparts.safeTag().toHost(), own = picparts.entOwners(2).toHost();
  std::vector<long> gids = picparts.globalIds(2).toHost();
  std::vector<int> lids = picparts.rankLocalIndex(2).toHost(), offs = picparts.nentsOffsets(2).toHost();
  std::vector<int> cai = picparts.commArrayIndex(2).toHost();
  CHECK((int)safe.size() == nel && (int)own.size() == nel && (int)gids.size() == nel && (int)offs.size() == NR + 1);
  CHECK(picparts.numBuffers(2) == (int)picparts.bufferedRanks(2).size() + 1);
  int ncore = 0, nsafe = 0;
  bool gid_ok = true, perm_ok = true;
  std::vector<char> seen(nel, 0);
  for (int e = 0; e < nel; ++e) {
    ncore += own[e] == 0;
    nsafe += safe[e] != 0;
    if (own[e] == 0 && !safe[e]) ++fails;                       // the core is always safe
    gid_ok = gid_ok && gids[e] == (long)offs[own[e]] + lids[e]; // owner-major global ids (part_construct.cpp:335-374)
    if (cai[e] < 0 || cai[e] >= nel || seen[cai[e]]) perm_ok = false; else seen[cai[e]] = 1;
  }
  CHECK(ncore == ne / NR && nsafe > ncore && nsafe < nel && gid_ok && perm_ok);
  CHECK((int)picparts.entOwners(0).size() == picparts.nents(0));

  // ---------------------------------------------------------------- pumipic::write / pumipic::read
  // (pumipic_mesh.hpp:147-151; test/test_file.cpp): every rank's PICpart written under one prefix,
  // rank 0's read back into an empty Mesh that owns its record
  {
    char tmpl[] = "/tmp/pp_mirror_XXXXXX";
    CHECK(mkdtemp(tmpl) != nullptr);
    const std::string prefix = std::string(tmpl) + "/plate";
    p::write(picparts, prefix.c_str());
    for (int r = 1; r < NR; ++r) {
      pp_host_picpart* other = nullptr;
      p::pp_check(pp_host_picpart_build(full, owner.data(), NR, r, 1, 1, -1, 2, &other), "picpart_build");
      p::pp_check(pp_host_picpart_write(other, prefix.c_str()), "picpart_write");
      pp_host_picpart_destroy(other);
    }
    pp_host_picpart* again = nullptr;
    p::pp_check(pp_host_picpart_read(prefix.c_str(), NR, 0, &again), "picpart_read");
    p::Mesh back;
    back.adopt(again, nullptr, true);
    CHECK(back.dim() == 2 && back.nelems() == nel && back.nents(0) == picparts.nents(0) && !back.isFullMesh());
    CHECK(back.safeTag().toHost() == safe && back.entOwners(2).toHost() == own);
    CHECK(back.globalIds(2).toHost() == gids && back.commArrayIndex(2).toHost() == cai);
    CHECK(back.bufferedRanks(2) == picparts.bufferedRanks(2));
    // Mesh(Input&) and the two owner-vector constructors on one rank (no communicator): the PICpart is
    // the whole mesh, everything is owned by rank 0 and safe
    std::vector<int> zero((size_t)ne, 0);
    p::Input in(full, p::Input::PARTITION, zero, p::Input::getMethod("bfs"), p::Input::getMethod("Full"));
    CHECK(in.bridge_dim == 0 && in.bufferBFSLayers == 3 && in.safeBFSLayers == 1 && in.getRule() == p::Input::PARTITION);
    CHECK(p::Input::getMethod("minimum") == p::Input::MINIMUM && p::Input::getMethod("x") == p::Input::INVALID);
    in.bridge_dim = 1;
    p::Mesh m_in(in);
    p::Mesh m_full(full, zero, nullptr);
    p::Mesh m_bfs(full, zero, 3, 1, nullptr);
    for (p::Mesh* mm : {&m_in, &m_full, &m_bfs}) {
      CHECK(mm->nelems() == ne && mm->numBuffers(2) == 1);
      std::vector<int> sf = mm->safeTag().toHost(), ow = mm->entOwners(2).toHost();
      int bad = 0;
      for (int e = 0; e < ne; ++e) bad += (sf[e] == 0) + (ow[e] != 0);
      CHECK(bad == 0);
    }
    CHECK(m_full.isFullMesh());
    bool threw = false;
    try { p::Mesh m_bad(full, zero, 1, 2, nullptr); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
  }

  // ---------------------------------------------------------------- structure with initial data (getMemberView)
  const int ppe_v = 20;
  std::vector<int> ppe(nel), pel;
  for (int e = 0; e < nel; ++e) { ppe[e] = safe[e] ? ppe_v + (e % 3) : 0; for (int k = 0; k < ppe[e]; ++k) pel.push_back(e); }
  const int np = (int)pel.size();
  p::MemberTypeViews info = p::createMemberViews<Particle>(np);
  auto ids_v = p::getMemberView<Particle, 0>(info);
  auto vec_v = p::getMemberView<Particle, 1>(info);
  auto int_v = p::getMemberView<Particle, 2>(info);
  CHECK(ids_v.size() == np && decltype(vec_v)::ncomp == 3);
  std::vector<int> h_ids(np), h_int(np);
  std::vector<double> h_vec(3 * (size_t)np);
  for (int i = 0; i < np; ++i) {
    h_ids[i] = i; h_int[i] = pel[i] * 7;
    for (int c = 0; c < 3; ++c) h_vec[(size_t)c * np + i] = pel[i] * (c + 1);   // test_structure.cpp:337-348
  }
  ids_v.fromHost(h_ids); vec_v.fromHost(h_vec); int_v.fromHost(h_int);
  CHECK(vec_v.toHost() == h_vec);
  PS::kkLidView ppe_d(ppe), pel_d(pel);
  PS::kkGidView gids_d(gids);
  p::TeamPolicy policy(10000, 32);
  PS* ptcls = new p::SellCSigma<Particle>(policy, INT_MAX, 1024, nel, np, ppe_d, gids_d, pel_d, info);
  p::destroyViews<Particle>(info);
  CHECK(ptcls->nPtcls() == np && ptcls->nElems() == nel && ptcls->capacity() >= np);
  {
    // the input-class constructors of the flat structures (dps/dps_input.hpp, cabm/cabm_input.hpp):
    // capacity = ceil(ceil(np / 32) * (1 + extra_padding)) * 32 (dps.hpp:129-132)
    p::DPS_Input<Particle> din(policy, nel, np, ppe_d, gids_d);
    din.extra_padding = 0.25;
    din.name = "dps_from_input";
    p::DPS<Particle> dps(din);
    CHECK(dps.nPtcls() == np && dps.capacity() == (int)std::ceil(std::ceil(np / 32.0) * 1.25) * 32);
    p::CabM_Input<Particle> cin(policy, nel, np, ppe_d, gids_d);
    p::CabM<Particle> cabm(cin);
    CHECK(cabm.nPtcls() == np && cabm.nElems() == nel);
    p::CSR_Input<Particle> csr_in(policy, nel, np, ppe_d, gids_d);
    CHECK(csr_in.padding_amount == 1.05 && !csr_in.always_realloc);
  }

  // every particle sits in the row of its element with its data (test_structure.cpp:326-351)
  const int cap = ptcls->capacity();
  p::View<int> bad(1, 0), slot_elem((size_t)cap, -1);
  {
    auto vec = ptcls->get<1>();
    auto tag = ptcls->get<2>();
    int* bad_p = bad.data();          // raw pointers: a lambda that crosses to the device captures PODs
    int* se_p = slot_elem.data();
    auto check = PS_LAMBDA(const int e, const int s, const bool mask) {
      if (mask) {
        se_p[s] = e;
        for (int c = 0; c < 3; ++c) if (vec(s, c) != (double)(e * (c + 1))) atomicAdd(bad_p, 1);
        if (tag(s) != e * 7) atomicAdd(bad_p, 1);
      }
    };
    p::parallel_for(ptcls, check, "check components");
    cudaDeviceSynchronize();
    CHECK(bad.toHost()[0] == 0);
  }

  // ---------------------------------------------------------------- getPIDs (testPIDs)
  {
    PS::kkLidView pids, offsets;
    ptcls->getPIDs(pids, offsets);
    std::vector<int> hp = pids.toHost(), ho = offsets.toHost(), se = slot_elem.toHost();
    CHECK((int)hp.size() == np && (int)ho.size() == nel + 1 && ho[0] == 0 && ho[nel] == np);
    int wrong = 0;
    std::vector<char> used(cap, 0);
    for (int i = 0; i < np; ++i) {
      const int s = hp[i];
      if (s < 0 || s >= cap || used[s] || se[s] < 0) { ++wrong; continue; }
      used[s] = 1;
      if (i < ho[se[s]] || i >= ho[se[s] + 1]) ++wrong;
    }
    for (int e = 0; e < nel; ++e) if (ho[e + 1] - ho[e] != ppe[e]) ++wrong;
    CHECK(wrong == 0);
  }

  // ---------------------------------------------------------------- setUnsafeProcs
  PS::kkLidView new_elems, new_procs;
  p::setUnsafeProcs(picparts, ptcls, slot_elem, new_elems, new_procs);
  cudaDeviceSynchronize();
  {
    std::vector<int> hne = new_elems.toHost(), hnp = new_procs.toHost(), se = slot_elem.toHost();
    int wrong = 0;
    for (int s = 0; s < cap; ++s) {
      if (hne[s] != se[s]) ++wrong;
      const int want = (se[s] >= 0 && !safe[se[s]]) ? own[se[s]] : 0;
      if (hnp[s] != want) ++wrong;
    }
    CHECK(wrong == 0);
  }

  // ---------------------------------------------------------------- trace_particle_through_mesh
  // stock handler: same arrays as the fused search_mesh; a user handler is called once per iteration
  {
    typedef p::Segment<double[3]> Seg3;
    const pp_host_mesh* pm = pp_host_picpart_mesh(rec);
    const double* pc = pp_host_mesh_coords(pm);
    const int32_t* pev = pp_host_mesh_ent2verts(pm, 2);
    std::vector<int> se = slot_elem.toHost();
    std::vector<double> hx(3 * (size_t)cap, 0.0), ht(3 * (size_t)cap, 0.0);
    for (int s = 0; s < cap; ++s) {
      if (se[s] < 0) continue;
      double cx = 0, cy = 0;
      for (int k = 0; k < 3; ++k) { cx += pc[2 * pev[3 * se[s] + k]]; cy += pc[2 * pev[3 * se[s] + k] + 1]; }
      hx[s] = cx / 3; hx[cap + s] = cy / 3;
      ht[s] = hx[s] + 0.45 + 0.001 * (s % 7); ht[cap + s] = hx[cap + s] + 0.2;   // some leave through x = 1
    }
    p::View<double> xv(hx), xtv(ht);
    Seg3 xs(xv.data(), cap), xts(xtv.data(), cap);
    auto pids = ptcls->get<0>();
    for (int mode = 0; mode < 2; ++mode) {          // 0: BCC walk, 1: edge intersection with wall points
      const bool req = mode == 1;
      p::View<int> ids_a, faces_a, ids_b, faces_b;
      p::View<double> pts_a, pts_b;
      const bool fa = p::search_mesh(picparts, ptcls, xs, xts, pids, ids_a, req, faces_a, pts_a, 500);
      struct Counting {
        p::RemoveParticleOnGeometricModelExit<Particle, Seg3> stock;
        int calls;
        void operator()(p::Mesh& m, PS* ps, p::View<int>& e, p::View<int>& f, p::View<int>& le, p::View<double>& xp,
                        p::View<int>& d, Seg3 a, Seg3 b) { ++calls; stock(m, ps, e, f, le, xp, d, a, b); }
      } handler{p::RemoveParticleOnGeometricModelExit<Particle, Seg3>(picparts, req), 0};
      const bool fb = p::trace_particle_through_mesh(picparts, ptcls, xs, xts, pids, ids_b, req, faces_b, pts_b,
                                                     500, false, handler);
      cudaDeviceSynchronize();
      std::vector<int> ia = ids_a.toHost(), ib = ids_b.toHost();
      int left = 0, stayed = 0;
      for (int s = 0; s < cap; ++s) if (se[s] >= 0) { left += ia[s] == -1; stayed += ia[s] == se[s]; }
      CHECK(fa && fb && ia == ib && handler.calls >= 3);
      CHECK(stayed < np / 2);                       // the push crosses several elements
      if (req) {
        std::vector<int> fA = faces_a.toHost(), fB = faces_b.toHost();
        int walls = 0;
        for (int s = 0; s < cap; ++s) walls += se[s] >= 0 && fA[s] >= 0;
        CHECK(fA == fB && pts_a.toHost() == pts_b.toHost() && walls > 0);
      } else {
        CHECK(left > 0 && left < np);
      }
    }
  }

  // ---------------------------------------------------------------- ParticleBalancer as rank 0 of 4, peers empty
  {
    p::ParticleBalancer balancer(picparts);
    picparts.setPtclBalancer(&balancer);
    CHECK(picparts.ptclBalancer() == &balancer && balancer.getSbarIDs(picparts).size() == (size_t)nel);
    p::printPtclImb(ptcls);                                     // pumipic_lb.hpp:379-398, one rank
    CHECK(picparts.mesh() == pp_host_picpart_mesh(rec));
    balancer.addWeights(picparts, ptcls, new_elems, new_procs);
    balancer.balance(picparts, 1.05);
    PS::kkLidView procs2(new_procs.toHost());
    balancer.selectParticles(picparts, ptcls, new_elems, procs2);
    cudaDeviceSynchronize();
    std::vector<int> before = new_procs.toHost(), after = procs2.toHost(), se = slot_elem.toHost();
    std::vector<int> sb = balancer.getSbarIDs(picparts).toHost();
    int32_t nsb = 0, mx = 0;
    const int32_t *sid, *soff, *sparts;
    p::pp_check(pp_host_picpart_sbars(rec, &nsb, &sid, &soff, &sparts, &mx), "sbars");
    int moved = 0, wrong = 0;
    std::vector<int> per_rank(NR, 0);
    for (int s = 0; s < cap; ++s) {
      if (se[s] < 0) { if (after[s] != before[s]) ++wrong; continue; }
      ++per_rank[after[s]];
      if (after[s] == before[s]) continue;
      ++moved;
      bool shares = false;   // the target belongs to the sbar of the particle's element
      for (int i = 0; i < nsb; ++i)
        if (sid[i] == sb[se[s]])
          for (int j = soff[i]; j < soff[i + 1]; ++j) shares = shares || sparts[j] == after[s];
      if (!shares || after[s] == 0) ++wrong;
    }
    CHECK(wrong == 0 && moved > 0 && per_rank[0] < np && per_rank[0] + per_rank[1] + per_rank[2] + per_rank[3] == np);
    // partition of particles per element (testBalanceArray)
    PS::kkLidView tgt = balancer.partition(picparts, ppe_d, 1.05);
    std::vector<int> ht = tgt.toHost();
    int out_of_range = 0, stay = 0;
    for (int v : ht) { out_of_range += v < 0 || v >= NR; stay += v == 0; }
    CHECK((int)ht.size() == np && out_of_range == 0 && stay > 0 && stay < np);
  }

  // ---------------------------------------------------------------- migrate_lb_ptcls on one process = rebuild
  {
    p::View<int> elems_next(slot_elem.toHost());
    p::migrate_lb_ptcls(picparts, ptcls, elems_next, 1.05f);
    CHECK(ptcls->nPtcls() == np);
  }

  // ---------------------------------------------------------------- PS_Comm_* on a one-rank communicator
  {
    pp_comm* comm = nullptr;
    p::pp_check(pp_comm_create(1, 0, nullptr, &comm), "comm");
    std::vector<double> hv(64);
    std::iota(hv.begin(), hv.end(), 1.0);
    p::View<double> a(hv), b((size_t)64, 0.0), c((size_t)64, 0.0), d((size_t)64, 0.0);
    CHECK(p::PS_Comm_Allreduce(a, b, 64, p::PS_SUM, comm) == 0 && b.toHost() == hv);
    CHECK(p::PS_Comm_Reduce(a, c, 64, p::PS_MAX, 0, comm) == 0 && c.toHost() == hv);
    CHECK(p::PS_Comm_Alltoall(a, 64, d, 64, comm) == 0 && d.toHost() == hv);
    p::PS_Request req;
    CHECK(p::PS_Comm_Wait(&req) == 0 && p::PS_Comm_Waitall(1, &req) == 0);   // nothing pending
    pp_comm_destroy(comm);
  }

  delete ptcls;
  pp_host_picpart_destroy(rec);
  pp_host_mesh_destroy(full);
  pp_host_free(co); pp_host_free(ev);
  if (fails) { fprintf(stderr, "%d check(s) failed\n", fails); return 1; }
  printf("MIRROR_API_OK\n");
  return 0;
}
