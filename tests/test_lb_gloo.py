"""CPU-only, 2 gloo ranks: the N>1 protocol of the particle balancer with the
device steps done in numpy -- every rank builds its PICpart and sbar table on its own, counts its
particles per own graph vertex, the global weight vector is summed by ONE all-reduce (here gloo, in
the product NCCL), every rank evaluates pp_host_lb_plan on its own and must arrive at the same plan,
takes its own sends from it, selects and "migrates" (an object all-gather).  testBalancePS of
test/test_lb.cpp:132-179: particles on rank 0 only, two rounds, imbalance <= 1.5 at the end, nothing
lost or duplicated, every particle still in an element its new rank holds safely."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        import torch
        import torch.distributed as dist
        from meshes import plate
        pp = importlib.import_module("pumi-pic_b200")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        m = plate(16)
        full = pp.HostMesh.from_elems(2, m.coords, m.elem2verts)
        cx = m.coords[m.elem2verts].mean(axis=1)[:, 0]
        owner = np.minimum((cx * world).astype(np.int32), world - 1)
        part = pp.Picpart.build(full, owner, world, rank, pp.BFS, pp.BFS, -1, 2)
        pm = part.mesh()
        sbar, own, safe = pm.tag(2, "sbar_id"), pm.tag(2, "ownership"), pm.tag(2, "safe")
        l2g = part.dim_info(2)["ent_l2g"]
        table, _ = part.sbars()
        # merge the tables like pp_balancer_create does: vertex -> (sbar, part), MAX over ranks
        nv = torch.tensor([max(g + len(p) for g, p in table.items())])
        dist.all_reduce(nv, op=dist.ReduceOp.MAX)
        nverts = int(nv)
        tab = torch.full((2, nverts), -1, dtype=torch.int64)
        for g, ps in table.items():
            for j, p in enumerate(ps):
                tab[0, g + j], tab[1, g + j] = g, p
        dist.all_reduce(tab, op=dist.ReduceOp.MAX)
        gtable = {}
        for v in range(nverts):
            if tab[0, v] >= 0:
                gtable.setdefault(int(tab[0, v]), []).append(int(tab[1, v]))
        gtable = {g: tuple(p) for g, p in gtable.items()}
        vof = {(g, p): g + j for g, ps in gtable.items() for j, p in enumerate(ps)}
        assert all(gtable[g] == ps for g, ps in table.items())
        # particles: (id, local element); 100 per core element on rank 0
        if rank == 0:
            core = np.flatnonzero(own == 0)
            elem = np.repeat(core, 100)
            pid = np.arange(elem.shape[0])
        else:
            elem, pid = np.zeros(0, np.int64), np.zeros(0, np.int64)
        total0 = torch.tensor([float(pid.shape[0])]); dist.all_reduce(total0)
        for rnd in range(2):
            w = torch.zeros(nverts + world, dtype=torch.float64)
            for g, c in zip(*np.unique(sbar[elem], return_counts=True)) if elem.shape[0] else []:
                if (int(g), rank) in vof:
                    w[vof[(int(g), rank)]] = float(c)
            dist.all_reduce(w)                                          # the one collective of the balancer
            sends, imb = pp.host_lb_plan(world, gtable, w[:nverts].numpy(), forced=w[nverts:].numpy(), tol=1.05)
            plans = [None] * world
            dist.all_gather_object(plans, sends)
            assert all(p == sends for p in plans), "ranks disagree on the plan"
            vpart = {v: gp for gp, v in vof.items()}
            new_rank = np.full(elem.shape[0], rank)
            for v, tgt, a in sends:
                g, p = vpart[v]
                if p != rank:
                    continue
                idx = np.flatnonzero((sbar[elem] == g) & (new_rank == rank))
                new_rank[idx[: int(np.ceil(a - 1e-9))]] = tgt
            out = [(int(i), int(l2g[e]), int(r)) for i, e, r in zip(pid, elem, new_rank) if r != rank]
            everything = [None] * world
            dist.all_gather_object(everything, out)
            keep = new_rank == rank
            g2l = {int(g): i for i, g in enumerate(l2g)}
            inc = [(i, g2l[ge]) for lst in everything for (i, ge, r) in lst if r == rank]
            pid = np.concatenate([pid[keep], np.asarray([i for i, _ in inc], np.int64)])
            elem = np.concatenate([elem[keep], np.asarray([e for _, e in inc], np.int64)])
            assert np.all(safe[elem] == 1), "a particle landed in an element this rank does not hold safely"
        counts = [None] * world
        dist.all_gather_object(counts, pid.tolist())
        allp = np.concatenate([np.asarray(c, np.int64) for c in counts])
        assert len(np.unique(allp)) == len(allp) == int(total0)
        per_rank = np.asarray([len(c) for c in counts], float)
        assert per_rank.max() / per_rank.mean() <= 1.5, per_rank
        q.put((rank, "ok", per_rank.tolist()))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "fail: %r\n%s" % (e, traceback.format_exc()), None))


def test_balancer_protocol_on_two_gloo_ranks():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 150
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
