"""CPU-only: host-side PICpart logic (pp_host_picpart_tags / pp_host_entity_owners) against a
numpy twin, and the N>1 protocol (owner/safe tags -> setUnsafeProcs -> exchange -> rebuild as set
semantics) on 2 gloo ranks with the oracle doing the geometry.  No GPU, no NCCL."""
import importlib
import os
import sys

import numpy as np
import pytest

from meshes import kuhn_cube, plate

HERE = os.path.dirname(os.path.abspath(__file__))


def _bfs_twin(mesh, owner, nranks, rank, buffer_method, safe_method, buffer_layers=3, safe_layers=1):
    FULL, BFS, MINIMUM, NONE = 0, 1, 2, 3
    if buffer_method == NONE:
        buffer_method = MINIMUM
    if buffer_method == MINIMUM:
        buffer_layers = 0
    if safe_method == MINIMUM:
        safe_layers = 0
    ev = mesh.elem2verts
    ne = mesh.nelems

    def layer(vis):
        touched = np.zeros(mesh.nverts, bool)
        touched[ev[vis.astype(bool)].ravel()] = True
        return (vis.astype(bool) | touched[ev].any(axis=1)).astype(np.int32)

    is_safe = np.full(ne, int(safe_method == FULL), np.int32)
    has_part = np.ones(nranks, np.int32)
    if (safe_method not in (NONE, FULL)) or buffer_method != FULL:
        vis = (owner == rank).astype(np.int32)
        safe = vis.copy()
        part = np.zeros(nranks, np.int32)
        part[rank] = 1
        for i in range(max(buffer_layers, safe_layers)):
            vis = layer(vis)
            if i == safe_layers - 1:
                safe = vis.copy()
            if i < buffer_layers:
                part[np.unique(owner[vis.astype(bool)])] = 1
        if safe_method in (BFS, MINIMUM):
            is_safe = safe
        if buffer_method in (BFS, MINIMUM):
            has_part = part
    if buffer_method == BFS and safe_method == FULL:
        vis = (1 - has_part[owner]).astype(np.int32)
        for i in range(safe_layers):
            vis = layer(vis)
        is_safe = ((vis == 0) | (owner == rank)).astype(np.int32)
    return is_safe, has_part


@pytest.mark.parametrize("meshname", ["cube", "plate"])
def test_picpart_tags_match_numpy_twin(meshname):
    pp = importlib.import_module("pumi-pic_b200")
    mesh = kuhn_cube(6) if meshname == "cube" else plate(24)
    cen = mesh.coords[mesh.elem2verts].mean(axis=1)
    nranks = 6
    owner = np.minimum((cen[:, 0] * nranks).astype(np.int32), nranks - 1)
    for bm in range(4):
        for sm in range(4):
            for rank in (0, 2, 5):
                got = pp.host_picpart_tags(mesh.dim, mesh.nverts, mesh.elem2verts, owner, nranks, rank,
                                           bm, sm, 3, 1)
                want = _bfs_twin(mesh, owner, nranks, rank, bm, sm)
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (bm, sm, rank)
                if sm != 3:
                    assert got[0][owner == rank].all()      # the core is always safe
    # defineOwners: vertex owner = min owner of adjacent elements
    vo = pp.host_entity_owners(mesh.nverts, mesh.elem2verts, owner, nranks)
    want = np.full(mesh.nverts, nranks, np.int32)
    np.minimum.at(want, mesh.elem2verts.ravel(), np.repeat(owner, mesh.dim + 1))
    assert np.array_equal(vo, want)


def _gloo_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                          WORLD_SIZE=str(world))
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        import torch.distributed as dist
        import oracle_api as orc
        import ptcl_init as pi
        pp = importlib.import_module("pumi-pic_b200")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        mesh = kuhn_cube(6)
        ne = mesh.nelems
        cen = mesh.coords[mesh.elem2verts].mean(axis=1)
        owner = np.minimum((cen[:, 0] * world).astype(np.int32), world - 1)
        safe, part = pp.host_picpart_tags(3, mesh.nverts, mesh.elem2verts, owner, world, rank)
        vo = pp.host_entity_owners(mesh.nverts, mesh.elem2verts, owner, world)
        # test_comm_array.cpp invariants: every vertex has exactly one owner over all ranks
        owned = [None] * world
        dist.all_gather_object(owned, (vo == rank).astype(np.int32))
        assert np.array_equal(sum(owned), np.ones(mesh.nverts, np.int32))
        tags = [None] * world
        dist.all_gather_object(tags, (safe, part))
        for r, (s, p) in enumerate(tags):
            assert s[owner == r].all() and p.all()
        # PIC loop with the oracle as the per-rank engine and a python exchange
        om = orc.OracleMesh(mesh)
        nptcl = 6000
        ppe = pi.even_ppe(ne, nptcl)
        se_g = np.repeat(np.arange(ne, dtype=np.int32), ppe)
        X, D = pi.init3d_internal(mesh, se_g, np.ones(nptcl, np.uint8))
        d = pi.push_distance(mesh) * 2.0
        mine = owner[se_g] == rank
        pid = np.nonzero(mine)[0]
        x = X[:, mine].copy(); dr = D[:, mine].copy(); el = se_g[mine].copy()
        Xo = X.copy(); ids_o = None
        for it in range(5):
            n = len(pid)
            t = x + d * dr
            _, ids, _, _, _ = om.search_mesh(el, np.ones(n, np.uint8), x, t)
            x = t
            ne_, proc = orc.set_unsafe_procs(np.ones(n, np.uint8), ids, safe, owner, rank)
            keep = (ids >= 0) & (proc == rank)
            out = [None] * world
            for p in range(world):
                sel = (ids >= 0) & (proc == p) & (p != rank)
                out[p] = (pid[sel], x[:, sel], dr[:, sel], ids[sel])
            inc = [None] * world
            dist.all_to_all_object_list = getattr(dist, "all_to_all_object_list", None)
            gathered = [None] * world
            dist.all_gather_object(gathered, out)
            inc = [gathered[p][rank] for p in range(world) if p != rank]
            pid = np.concatenate([pid[keep]] + [a[0] for a in inc])
            x = np.concatenate([x[:, keep]] + [a[1] for a in inc], axis=1)
            dr = np.concatenate([dr[:, keep]] + [a[2] for a in inc], axis=1)
            el = np.concatenate([ids[keep]] + [a[3] for a in inc]).astype(np.int32)
            To = Xo + d * D
            _, ids_o, _, _, _ = om.search_mesh(se_g if ids_o is None else np.maximum(ids_o, 0),
                                               np.ones(nptcl, np.uint8) if ids_o is None
                                               else (ids_o >= 0).astype(np.uint8), Xo, To)
            Xo = To
            allp = [None] * world
            dist.all_gather_object(allp, (pid, el))
            ids_all = np.concatenate([a[0] for a in allp]); el_all = np.concatenate([a[1] for a in allp])
            assert len(np.unique(ids_all)) == len(ids_all)
            alive = np.nonzero(ids_o >= 0)[0]
            order = np.argsort(ids_all)
            assert np.array_equal(ids_all[order], alive)
            assert np.array_equal(el_all[order], ids_o[alive])
            assert np.all((safe[el] == 1) | (owner[el] == rank))
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def test_two_rank_protocol_on_gloo():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res


def test_picpart_extract_matches_numpy_twin():
    """constructPICPart's sub-mesh (part_construct.cpp:116-262): kept elements and their vertices in
    full-mesh order, local connectivity, gathered coordinates; slabs so that far cores are dropped."""
    pp = importlib.import_module("pumi-pic_b200")
    for mesh in (kuhn_cube(8), plate(32)):
        cen = mesh.coords[mesh.elem2verts].mean(axis=1)
        nranks = 4
        owner = np.minimum((cen[:, 0] * nranks).astype(np.int32), nranks - 1)
        for rank in range(nranks):
            safe, part = pp.host_picpart_tags(mesh.dim, mesh.nverts, mesh.elem2verts, owner, nranks, rank,
                                              pp.api.BFS, pp.api.BFS, 1, 1)
            assert part[rank] == 1 and part.sum() < nranks          # partially buffered
            el2g, vl2g, evl, col = pp.host_picpart_extract(mesh.dim, mesh.coords, mesh.elem2verts, owner,
                                                           nranks, part)
            keep = part[owner].astype(bool)
            assert np.array_equal(el2g, np.nonzero(keep)[0])
            assert np.array_equal(vl2g, np.unique(mesh.elem2verts[keep]))
            assert np.array_equal(vl2g[evl], mesh.elem2verts[keep])
            assert np.array_equal(col, mesh.coords[vl2g])
            # the local mesh is a valid simplicial mesh for the search: sides can be derived
            e2s, s2v = pp.host_derive_sides(mesh.dim, evl)
            assert e2s.shape == evl.shape and s2v.max() < len(vl2g)
            # every safe element is kept, and the kept set is core + whole neighbouring cores
            assert keep[safe.astype(bool)].all()
