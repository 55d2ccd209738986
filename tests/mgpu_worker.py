"""Multi-GPU worker (launched with torchrun by tests/test_multigpu.py or by hand under
`gpurun --gpus N`).  One process per GPU, NCCL through the C ABI's pp_comm.

Checks, following particle_structs/test/test_migrate.cpp and test/test_comm_array.cpp:
  1. migrate: send right and back, 5 % to rank 0, empty and refill -- by particle id
  2. comm-array reductions: SUM of ones == nranks, MIN of owners, BCAST owner's value
  3. a full PIC loop (fused push+search -> setUnsafeProcs -> migrate) on a block-partitioned
     Kuhn cube: the union over ranks of (particle id -> element) must equal the serial CPU oracle
  4. the same loop and the comm-array reductions on partially buffered PICparts (sub-meshes)
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_api as orc          # noqa: E402
import ptcl_init as pi            # noqa: E402
from meshes import kuhn_cube      # noqa: E402

TYPES = [(np.int32, 1), (np.float64, 3), (np.int32, 1)]
PIC = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]


SHARED_GPU = os.environ.get("MGPU_SHARED_GPU") == "1"   # all ranks on cuda:0, no NCCL (single-GPU box)
HOSTED = SHARED_GPU or os.environ.get("MGPU_HOSTED") == "1"


def make_comm(P):
    """NCCL id over torch.distributed (default), or the hosted bootstrap (pp_comm_create_hosted) with
    torch.distributed's all-gather as the application's callback -- without NCCL when the ranks share a GPU."""
    if HOSTED:
        return P.Comm(hosted=True, nccl=not SHARED_GPU)
    return P.Comm()


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def gather_np(a):
    """all ranks' numpy arrays, concatenated (variable length)."""
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, a)
    return objs


def ids_and_elems(ps):
    se, m = ps.slot_elem_and_mask()
    m = m.astype(bool)
    return ps.get(0).cpu().numpy()[0, :ps.capacity][m], se[m]


def test_migrate(P, comm, rank, R):
    ne, npr = 200, 5000
    rng = np.random.default_rng(10 + rank)
    ppe = rng.multinomial(npr, np.ones(ne) / ne).astype(np.int32)
    gids = (np.arange(ne, dtype=np.int64) * 3 + 7)                 # non-trivial global ids
    pel = np.repeat(np.arange(ne, dtype=np.int32), ppe)
    ids = (np.arange(npr, dtype=np.int32) + rank * 1000000)
    info = [ids.reshape(1, -1), rng.random((3, npr)), (ids % 17).reshape(1, -1).astype(np.int32)]
    payload = {int(i): info[1][:, k].copy() for k, i in enumerate(ids)}
    for kind in (P.capi.PP_PS_SCS, P.capi.PP_PS_CSR, P.capi.PP_PS_DPS):
        if kind == P.capi.PP_PS_CSR:
            ps = P.ParticleStructure(kind, TYPES, np.zeros(ne, np.int32), elem_gids=gids)
            ps.rebuild(torch.full((max(ps.capacity, 1),), -1, dtype=torch.int32, device="cuda"),
                       dev(pel), [dev(a) for a in info])
        else:
            ps = P.ParticleStructure(kind, TYPES, ppe, elem_gids=gids, particle_elements=pel,
                                     particle_info=info)
        all_payload = {}
        for d in gather_np(payload):
            all_payload.update(d)
        # 1. send every 3rd particle to the right neighbour, into element (e+1)%ne
        for step, shift in ((0, 1), (1, -1)):
            se, m = ps.slot_elem_and_mask()
            m = m.astype(bool)
            cap = ps.capacity
            pid = ps.get(0).cpu().numpy()[0, :cap]
            new_elem = np.where(m, (se + 1) % ne, -1).astype(np.int32)
            moving = m & (pid % 3 == 0)
            new_proc = np.where(moving, (rank + shift) % R, rank).astype(np.int32)
            before = dict(zip(pid[m].tolist(), zip(new_elem[m].tolist(), new_proc[m].tolist())))
            ne_d = dev(new_elem)
            sent, recv = P.migrate(ps, comm, ne_d, dev(new_proc))
            assert sent == (int(moving.sum()) if R > 1 else 0)
            got_ids, got_elems = ids_and_elems(ps)
            want = {}
            for d in gather_np(before):
                want.update(d)
            mine = {i: e for i, (e, p) in want.items() if p == rank}
            assert dict(zip(got_ids.tolist(), got_elems.tolist())) == mine, "migrate destination mismatch"
            vec = ps.get(1).cpu().numpy()[:, :ps.capacity]
            se2, m2 = ps.slot_elem_and_mask()
            for k in np.nonzero(m2)[0][:: max(1, m2.sum() // 200)]:
                i = int(ps.get(0)[0, k])
                assert np.array_equal(vec[:, k], all_payload[i]), "payload corrupted in flight"
        tot = sum(len(x) for x in gather_np(ids_and_elems(ps)[0]))
        assert tot == npr * R
        # 2. 5 % of the particles to rank 0 (test_migrate.cpp "sendToOne")
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        pid = ps.get(0).cpu().numpy()[0, :ps.capacity]
        new_proc = np.where(m & (pid % 20 == 0), 0, rank).astype(np.int32)
        new_elem = np.where(m, se, -1).astype(np.int32)
        P.migrate(ps, comm, dev(new_elem), dev(new_proc))
        counts = [len(x) for x in gather_np(ids_and_elems(ps)[0])]
        assert sum(counts) == npr * R and (R == 1 or counts[0] > npr)
        # 3. empty every rank but 0, then refill with new particles through migrate
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        new_elem = np.where(m, se, -1).astype(np.int32)
        new_proc = np.zeros(ps.capacity, np.int32)
        P.migrate(ps, comm, dev(new_elem), dev(new_proc))
        counts = [len(x) for x in gather_np(ids_and_elems(ps)[0])]
        assert counts[0] == npr * R and all(c == 0 for c in counts[1:])
        n_new = 321
        nel = (np.arange(n_new) % ne).astype(np.int32)
        ninfo = [dev((np.arange(n_new, dtype=np.int32) + 5000000 + rank * 1000).reshape(1, -1)),
                 torch.ones((3, n_new), dtype=torch.float64, device="cuda"),
                 torch.zeros((1, n_new), dtype=torch.int32, device="cuda")]
        cap = max(ps.capacity, 1)
        keep = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
        if rank == 0:
            se, m = ps.slot_elem_and_mask()
            keep = dev(np.where(m.astype(bool), se, -1).astype(np.int32))
        P.migrate(ps, comm, keep, torch.full((cap,), rank, dtype=torch.int32, device="cuda"),
                  dev(nel), ninfo)
        counts = [len(x) for x in gather_np(ids_and_elems(ps)[0])]
        assert counts[0] == npr * R + n_new and all(c == n_new for c in counts[1:])
    if rank == 0:
        print("migrate scenarios ok on %d ranks" % R)


def test_comm_array(P, comm, rank, R):
    n, nv = 1000, 2
    ones = torch.ones(n * nv, dtype=torch.float64, device="cuda")
    comm.array_reduce(ones, n, nv, P.capi.PP_SUM)
    assert float(ones.min()) == R and float(ones.max()) == R          # test_comm_array.cpp:119-204
    owner = (np.arange(n) % R).astype(np.int32)
    t = torch.full((n,), rank, dtype=torch.int32, device="cuda")
    comm.array_reduce(t, n, 1, P.capi.PP_MIN)
    assert int(t.max()) == 0
    t = torch.full((n,), rank, dtype=torch.int32, device="cuda")
    comm.array_reduce(t, n, 1, P.capi.PP_MAX)
    assert int(t.min()) == R - 1
    val = (torch.arange(n * nv, dtype=torch.float64, device="cuda") + 1000 * rank)
    comm.array_reduce(val, n, nv, P.capi.PP_BCAST, dev(owner))
    want = np.arange(n * nv) + 1000 * np.repeat(owner, nv)
    assert np.array_equal(val.cpu().numpy(), want)
    a = torch.arange(R * 3, dtype=torch.int32, device="cuda") + 100 * rank
    b = torch.empty_like(a)
    if not SHARED_GPU:
        comm.alltoall(a, b)
        want = np.concatenate([np.arange(3) + 3 * rank + 100 * p for p in range(R)])
        assert np.array_equal(b.cpu().numpy(), want)
    elif R > 1:
        # a communicator without NCCL says so instead of crashing
        try:
            comm.alltoall(a, b)
            raise AssertionError("alltoall on a communicator without NCCL must fail")
        except P.PumipicError as e:
            assert "no NCCL transport" in str(e)
        # PS_Comm_Allreduce runs over the peer-memory window there
        t = dev(np.arange(5, dtype=np.float64) + rank)
        comm.allreduce(t)
        assert np.array_equal(t.cpu().numpy(), R * np.arange(5.0) + R * (R - 1) / 2)
    # a larger array (the window grows, collectively), random doubles: the sum in ascending rank order,
    # bit for bit, on every rank
    for n2 in (50000, 777777):
        rng = np.random.default_rng(100 + rank)
        mine = rng.standard_normal(n2)
        got = comm.array_reduce(dev(mine), n2, 1, P.capi.PP_SUM).cpu().numpy()
        parts = gather_np(mine)
        want = parts[0].copy()
        for q in range(1, R):
            want = want + parts[q]
        assert np.array_equal(got, want), "rank-ordered sum differs"
        f = dev(mine.astype(np.float32))
        got32 = comm.array_reduce(f, n2, 1, P.capi.PP_MAX).cpu().numpy()
        assert np.array_equal(got32, np.max(np.stack([q.astype(np.float32) for q in parts]), axis=0))
    if rank == 0:
        print("comm arrays ok on %d ranks" % R)


def pic_loop_parity(P, comm, rank, R, steps=6, n=8, nptcl=40000, fuse_update=False):
    """The multi-rank PIC loop (fused push+search -> updatePtclPositions -> setUnsafeProcs -> migrate)
    on a block-partitioned Kuhn cube against the serial CPU oracle on the whole particle set, by
    particle id.  Returns mismatch counts (all ranks return the same dict); also called by bench.py's
    `picstep` leg as the in-run checker."""
    mesh = kuhn_cube(n)
    ne = mesh.nelems
    # block partition along x (and y for R >= 4): element centroid decides the owner
    cen = mesh.coords[mesh.elem2verts].mean(axis=1)
    bx = 2 if R >= 2 else 1
    by = 2 if R >= 4 else 1
    bz = 2 if R >= 8 else 1
    owner = ((cen[:, 0] * bx).astype(int).clip(0, bx - 1)
             + bx * ((cen[:, 1] * by).astype(int).clip(0, by - 1)
                     + by * (cen[:, 2] * bz).astype(int).clip(0, bz - 1))).astype(np.int32) % R
    safe, part = P.host_picpart_tags(3, mesh.nverts, mesh.elem2verts, owner, R, rank)
    assert part.all() and safe[owner == rank].all()
    gm = P.Mesh(3, mesh.coords, mesh.elem2verts, mesh.elem2sides, mesh.side2verts, mesh.class_id)
    gm.set_picpart(safe, owner, rank)
    # global particle set (identical on every rank), each rank keeps those in its own core
    ppe_g = pi.even_ppe(ne, nptcl)
    slot_elem_g = np.repeat(np.arange(ne, dtype=np.int32), ppe_g)
    mask_g = np.ones(nptcl, np.uint8)
    X, D = pi.init3d_internal(mesh, slot_elem_g, mask_g)
    dist_push = pi.push_distance(mesh) * 2.5
    mine = owner[slot_elem_g] == rank
    pel = slot_elem_g[mine]
    info = [X[:, mine], np.zeros((3, mine.sum())), np.nonzero(mine)[0].astype(np.int32).reshape(1, -1),
            D[:, mine]]
    ppe = np.bincount(pel, minlength=ne).astype(np.int32)
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, PIC, ppe, elem_gids=np.arange(ne, dtype=np.int64),
                             particle_elements=pel, particle_info=info)
    # serial oracle on the whole set
    om = orc.OracleMesh(mesh)
    Xo = X.copy(); ids_o = None
    out = {"ranks": R, "steps": steps, "particles": nptcl, "duplicated": 0, "missing_or_extra": 0,
           "mismatch_element": 0, "mismatch_position": 0, "on_unsafe_rank": 0, "migrated": 0}
    for it in range(steps):
        cap = ps.capacity
        x, tg, pid, dr = ps.get(0), ps.get(1), ps.get(2), ps.get(3)
        ids = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
        P.push_direction_search(gm, ps, dr, dist_push, x, tg, ids, elem_ids_empty=True, from_orig=True)
        if fuse_update and it % 2 == 0:         # updatePtclPositions as a member remap of the rebuild
            ps.set_rebuild_remap([1, -1, 2, 3])
        else:
            P.update_positions(ps, x, tg)
        ne_d, np_d = P.set_unsafe_procs(gm, ps, ids)
        sent, _ = P.migrate(ps, comm, ne_d, np_d)
        # oracle step
        To = Xo + dist_push * D
        found, ids_o, _, _, st = om.search_mesh(slot_elem_g if ids_o is None else np.maximum(ids_o, 0),
                                                (mask_g if ids_o is None else (ids_o >= 0).astype(np.uint8)),
                                                Xo, To)
        Xo = To
        # compare by particle id
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        pids = ps.get(2).cpu().numpy()[0, :ps.capacity][m]
        xs = ps.get(0).cpu().numpy()[:, :ps.capacity][:, m]
        unsafe_here = int(np.sum(~((safe[se[m]] == 1) | (owner[se[m]] == rank))))
        got = {}
        parts = gather_np((pids, se[m], xs, unsafe_here, sent)) if R > 1 else [(pids, se[m], xs, unsafe_here, sent)]
        for d in parts:
            out["on_unsafe_rank"] += d[3]
            out["migrated"] += d[4]
            for i, e, xx in zip(d[0].tolist(), d[1].tolist(), d[2].T):
                if i in got:
                    out["duplicated"] += 1
                got[i] = (e, xx)
        alive = np.nonzero(ids_o >= 0)[0]
        out["missing_or_extra"] += len(set(got).symmetric_difference(alive.tolist()))
        for i in alive:
            g = got.get(int(i))
            if g is None:
                continue
            out["mismatch_element"] += int(g[0] != ids_o[i])
            out["mismatch_position"] += int(not np.array_equal(g[1], Xo[:, i]))
    out["alive_at_end"] = int(len(alive))
    out["mismatch"] = (out["duplicated"] + out["missing_or_extra"] + out["mismatch_element"]
                       + out["mismatch_position"] + out["on_unsafe_rank"])
    return out


def test_pic_loop(P, comm, rank, R, steps=6):
    r = pic_loop_parity(P, comm, rank, R, steps)
    assert r["mismatch"] == 0, r
    r2 = pic_loop_parity(P, comm, rank, R, steps, fuse_update=True)
    assert r2["mismatch"] == 0 and r2["alive_at_end"] == r["alive_at_end"], r2
    assert R == 1 or r["migrated"] > 0
    if rank == 0:
        print("PIC loop parity ok on %d ranks: %d of %d particles still in the domain, %d migrations"
              % (R, r["alive_at_end"], r["particles"], r["migrated"]))


def test_partial_picparts(P, comm, rank, R, steps=6):
    """Partially buffered PICparts (Input BFS buffer / BFS safe, part_construct.cpp:116-262): every
    rank holds only its core and the cores within 3 BFS layers, as a renumbered sub-mesh.  Checks
    (a) Mesh::reduceCommArray's owner fan-in / fan-out (pumipic_comm.cpp:249-439) through
    pp_comm_plan_* against numpy on the gathered copies, (b) the PIC loop with migration by global
    element id against the serial oracle on the full mesh, by particle id, bit-exact."""
    n = 16
    mesh = kuhn_cube(n)
    ne = mesh.nelems
    cen = mesh.coords[mesh.elem2verts].mean(axis=1)
    owner = np.minimum((cen[:, 0] * R).astype(np.int32), R - 1)          # slabs along x
    safe_g, part = P.host_picpart_tags(3, mesh.nverts, mesh.elem2verts, owner, R, rank,
                                       P.api.BFS, P.api.BFS, 3, 1)
    if R >= 4:
        assert part.sum() < R, "slabs 4 cells wide: 3 BFS layers must not reach the second neighbour"
    el2g, vl2g, evl, col = P.host_picpart_extract(3, mesh.coords, mesh.elem2verts, owner, R, part)
    e2s, s2v = P.host_derive_sides(3, evl)
    gm = P.Mesh(3, col, evl, e2s, s2v, np.ones(len(el2g), np.int32))
    gm.set_picpart(safe_g[el2g], owner[el2g], rank)
    # ---- (a) vertex comm array over the plan
    vo_g = P.host_entity_owners(mesh.nverts, mesh.elem2verts, owner, R)
    plan = comm.plan(vl2g.astype(np.int64) * 5 + 3, vo_g[vl2g])           # non-trivial global ids
    nv = 2
    holders = gather_np(vl2g)
    nh = np.zeros(mesh.nverts)
    for h in holders:
        nh[h] += 1
    ns, nr = plan.counts()
    assert ns == int((vo_g[vl2g] != rank).sum())
    assert nr == int((nh[vl2g][vo_g[vl2g] == rank] - 1).sum())
    base = (np.arange(len(vl2g) * nv) % 7) * 0.125
    for dt in (torch.float64, torch.int32):
        mine_v = (np.repeat(vl2g, nv) % 13 + 10.0 * rank + (base if dt == torch.float64 else 0))
        vals = gather_np(mine_v)
        for op, red in ((P.capi.PP_SUM, np.add), (P.capi.PP_MAX, np.maximum), (P.capi.PP_MIN, np.minimum)):
            want = np.full(mesh.nverts * nv, np.nan)
            for h, v in zip(holders, vals):        # ascending rank order, like the plan's merge
                idx = (np.repeat(h, nv) * nv + np.tile(np.arange(nv), len(h)))
                cur = want[idx]
                want[idx] = np.where(np.isnan(cur), v, red(cur, v))
            arr = torch.as_tensor(mine_v).to(dt).cuda()
            plan.reduce(arr, nv, op)
            idx = (np.repeat(vl2g, nv) * nv + np.tile(np.arange(nv), len(vl2g)))
            got = arr.cpu().numpy().astype(np.float64)
            if op == P.capi.PP_SUM and dt == torch.float64 and R > 2:
                # the owner adds its own value first, then the others in rank order
                assert np.allclose(got, want[idx], rtol=1e-14, atol=0)
            else:
                assert np.array_equal(got, want[idx]), (op, dt)
        arr = torch.as_tensor(mine_v).to(dt).cuda()
        plan.reduce(arr, nv, P.capi.PP_BCAST)
        own_val = np.repeat(vl2g, nv) % 13 + 10.0 * np.repeat(vo_g[vl2g], nv)
        got = arr.cpu().numpy().astype(np.float64)
        if dt == torch.float64:
            # the owner's `base` term is indexed by the owner's local numbering: fetch it from there
            want = np.full(mesh.nverts * nv, np.nan)
            for r_, (h, v) in enumerate(zip(holders, vals)):
                sel = np.repeat(vo_g[h] == r_, nv)
                idx_h = (np.repeat(h, nv) * nv + np.tile(np.arange(nv), len(h)))
                want[idx_h[sel]] = v[sel]
            assert np.array_equal(got, want[idx])
        else:
            assert np.array_equal(got, own_val)
    # ---- (b) PIC loop on the sub-mesh, particles migrate by global element id
    nptcl = 60000
    ppe_g = pi.even_ppe(ne, nptcl)
    slot_elem_g = np.repeat(np.arange(ne, dtype=np.int32), ppe_g)
    mask_g = np.ones(nptcl, np.uint8)
    X, D = pi.init3d_internal(mesh, slot_elem_g, mask_g)
    dist_push = pi.push_distance(mesh) * 2.5
    g2l = np.full(ne, -1, np.int32)
    g2l[el2g] = np.arange(len(el2g), dtype=np.int32)
    mine = owner[slot_elem_g] == rank
    pel = g2l[slot_elem_g[mine]]
    assert (pel >= 0).all()
    info = [X[:, mine], np.zeros((3, mine.sum())), np.nonzero(mine)[0].astype(np.int32).reshape(1, -1),
            D[:, mine]]
    ppe = np.bincount(pel, minlength=len(el2g)).astype(np.int32)
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, PIC, ppe, elem_gids=el2g.astype(np.int64),
                             particle_elements=pel, particle_info=info)
    om = orc.OracleMesh(mesh)
    Xo = X.copy(); ids_o = None
    migrated = 0
    safe_l, owner_l = safe_g[el2g], owner[el2g]
    for it in range(steps):
        cap = ps.capacity
        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        ids = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
        P.push_direction_search(gm, ps, dr, dist_push, x, tg, ids, elem_ids_empty=True, from_orig=True)
        P.update_positions(ps, x, tg)
        ne_d, np_d = P.set_unsafe_procs(gm, ps, ids)
        sent, recv = P.migrate(ps, comm, ne_d, np_d)
        migrated += sent
        To = Xo + dist_push * D
        found, ids_o, _, _, st = om.search_mesh(slot_elem_g if ids_o is None else np.maximum(ids_o, 0),
                                                (mask_g if ids_o is None else (ids_o >= 0).astype(np.uint8)),
                                                Xo, To)
        Xo = To
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        pids = ps.get(2).cpu().numpy()[0, :ps.capacity][m]
        xs = ps.get(0).cpu().numpy()[:, :ps.capacity][:, m]
        got = {}
        for d in gather_np((pids, el2g[se[m]], xs)):
            for i, e, xx in zip(d[0].tolist(), d[1].tolist(), d[2].T):
                assert i not in got, "particle %d lives on two ranks" % i
                got[i] = (e, xx)
        alive = np.nonzero(ids_o >= 0)[0]
        assert sorted(got) == alive.tolist(), "particle set differs from the serial oracle"
        for i in alive[:: max(1, len(alive) // 3000)]:
            assert got[i][0] == ids_o[i] and np.array_equal(got[i][1], Xo[:, i])
        assert np.all((safe_l[se[m]] == 1) | (owner_l[se[m]] == rank))
    tot = sum(gather_np(migrated))
    assert R == 1 or tot > 0
    if rank == 0:
        print("partial PICparts ok on %d ranks: %d local of %d elements, %d particles migrated"
              % (R, len(el2g), ne, tot))


def test_small_window(P, rank, R):
    """Peer-memory window with room for 100 particles per peer: a step that wants to send more keeps the
    rest (stats.deferred) with their new_process still set; repeating the migration drains them, and
    no particle is lost or duplicated."""
    comm = make_comm(P)
    ne, npr = 64, 3000
    ppe = np.full(ne, npr // ne, np.int32); ppe[: npr % ne] += 1
    pel = np.repeat(np.arange(ne, dtype=np.int32), ppe)
    ids = np.arange(npr, dtype=np.int32) + rank * 1000000
    info = [ids.reshape(1, -1), np.zeros((3, npr)), (ids % 17).reshape(1, -1).astype(np.int32)]
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, TYPES, ppe, elem_gids=np.arange(ne, dtype=np.int64),
                             particle_elements=pel, particle_info=info)
    comm.set_p2p_window(100 * 48)          # TYPES travels as 48-byte records (gid 8 + 8 + 24 + 8): 100 per peer
    total_sent = 0
    for it in range(40):
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        pid = ps.get(0).cpu().numpy()[0, :ps.capacity]
        here = m & (pid // 1000000 == rank)              # my original particles all go to the right
        new_elem = np.where(m, se, -1).astype(np.int32)
        new_proc = np.where(here, (rank + 1) % R, rank).astype(np.int32)
        st = P.capi.MigrateStats()
        P.capi.check(P.lib().pp_ps_migrate(ps.h, comm.h, P.api._ptr(dev(new_elem)), P.api._ptr(dev(new_proc)), 0,
                                           None, None, C.byref(st), P.api._stream()))
        total_sent += st.sent
        left = gather_np(int(here.sum()) - int(st.sent))
        if it == 0:
            assert comm.p2p_active and st.deferred > 0 and st.sent < here.sum()
        if sum(left) == 0:
            break
    assert total_sent == npr, (total_sent, npr)
    se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
    pid = ps.get(0).cpu().numpy()[0, :ps.capacity][m]
    allp = np.concatenate(gather_np(pid))
    assert len(allp) == R * npr and len(np.unique(allp)) == len(allp)
    assert np.all(pid // 1000000 == (rank - 1) % R)      # everything I hold came from my left neighbour
    if rank == 0:
        print("small window ok on %d ranks: drained in %d migrations" % (R, it + 1))


def test_balancer(P, comm, rank, R):
    """testBalancePS (test/test_lb.cpp:132-207): 100 particles per element on the even ranks only,
    two rounds of repartition + migrate; the imbalance must end <= 1.5 and no particle may be lost,
    duplicated or moved to another element."""
    from meshes import plate
    m = plate(16)
    full = P.HostMesh.from_elems(2, m.coords, m.elem2verts)
    cx = m.coords[m.elem2verts].mean(axis=1)[:, 0]
    owner = np.minimum((cx * R).astype(np.int32), R - 1)            # R stripes
    part = P.Picpart.build(full, owner, R, rank, P.BFS, P.FULL)
    pm = part.mesh()
    sbar, own, safe = pm.tag(2, "sbar_id"), pm.tag(2, "ownership"), pm.tag(2, "safe")
    gids = pm.tag(2, "gids").astype(np.int64)
    table, _ = part.sbars()
    bal = P.Balancer(R, rank, table, sbar.astype(np.int32), own.astype(np.int32), comm=comm)
    ne = pm.nents(2)
    ppe = np.where(own == rank, 100 if rank % 2 == 0 else 0, 0).astype(np.int32)
    npr = int(ppe.sum())
    pel = np.repeat(np.arange(ne, dtype=np.int32), ppe)
    ids = np.arange(npr, dtype=np.int32) + rank * 1000000
    info = [ids.reshape(1, -1), np.zeros((3, npr)), gids[pel].astype(np.int32).reshape(1, -1)]
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, TYPES, ppe, elem_gids=gids, particle_elements=pel,
                             particle_info=info)
    d_safe, d_own = dev(safe.astype(np.int32)), dev(own.astype(np.int32))
    before = gather_np(ps.nptcls)
    for _ in range(2):
        se, mk = ps.slot_elem_and_mask()
        d_se, d_mk = dev(se), dev(mk.astype(bool))
        new_elems = torch.where(d_mk, d_se, torch.full_like(d_se, -1))
        unsafe = d_mk & (d_safe[d_se.long()] == 0)
        new_procs = torch.where(unsafe, d_own[d_se.long()], torch.full_like(d_se, rank))
        bal.repartition(comm, ps, 1.05, new_elems, new_procs)
        P.migrate(ps, comm, new_elems, new_procs)
    after = gather_np(ps.nptcls)
    assert sum(after) == sum(before)
    imb = max(after) / (sum(after) / R)
    assert imb <= 1.5, (before, after)
    se, mk = ps.slot_elem_and_mask()
    mk = mk.astype(bool)
    pid = ps.get(0).cpu().numpy()[0, :ps.capacity][mk]
    home = ps.get(2).cpu().numpy()[0, :ps.capacity][mk]
    assert np.array_equal(gids[se[mk]].astype(np.int32), home)     # still in its element
    allp = np.concatenate(gather_np(pid))
    assert len(np.unique(allp)) == len(allp) == sum(before)
    if rank == 0:
        print("balancer ok on %d ranks: particles %s -> %s (imbalance %.3f)" % (R, before, after, imb))


def main():
    import faulthandler
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True)   # `timeout -s USR1` prints where a rank hangs
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    R = int(os.environ.get("WORLD_SIZE", "1"))
    if SHARED_GPU:
        local = 0
    torch.cuda.set_device(local)
    nccl = R > 1 and not SHARED_GPU
    dist.init_process_group("nccl" if nccl else "gloo", device_id=torch.device("cuda", local) if nccl else None)
    P = importlib.import_module("pumi-pic_b200")
    comm = make_comm(P)
    only = os.environ.get("MGPU_ONLY", "")       # e.g. MGPU_ONLY=balancer for one scenario
    if only in ("", "comm_array"):
        test_comm_array(P, comm, rank, R)
    if only in ("", "migrate"):
        test_migrate(P, comm, rank, R)
    if only in ("", "pic_loop"):
        test_pic_loop(P, comm, rank, R)
    if only in ("", "partial") and not SHARED_GPU:     # comm plans are NCCL send/recv
        test_partial_picparts(P, comm, rank, R)
    # One rank: repartition is a no-op (pumipic_lb.hpp:360-361)
    if R > 1 and only in ("", "balancer"):
        test_balancer(P, comm, rank, R)
    if R > 1 and only == "small_window":
        test_small_window(P, rank, R)
    if R > 1 and only != "comm_array" and os.environ.get("MGPU_EXPECT_P2P") == "1":
        assert comm.p2p_active or only == "small_window", "the peer-memory window was expected to be in use"
    dist.barrier()
    if rank == 0:
        print("MGPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a rank that fails must not sit in NCCL teardown while its peers wait in a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
