#!/usr/bin/env python
"""Weak-scaling bench of the FULL PIC step on N GPUs (BASELINE configs[4]).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_picstep.py [--cube-per-gpu 55] [--ppe 10] [--steps 10]

Global mesh: Kuhn cube split into bx*by*bz blocks (one PICpart core per GPU, ~6*cube^3 tets each),
every rank buffers the full mesh (Input::FULL, so the ghost reduction is one all-reduce,
pumipic_comm.cpp:234-247) and is safe within one BFS layer of its core (pumipic_input.cpp:103-110).
Step = fused push+search -> updatePtclPositions -> setUnsafeProcs -> migrate (NCCL all-to-all-v +
rebuild) -> vertex comm-array all-reduce (the gyroSync stand-in for a 3D run).  CUDA events per
phase; the step time is the max over ranks.  One JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube-per-gpu", type=int, default=55)
    ap.add_argument("--ppe", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--push-mult", type=float, default=3.0, help="push distance in units of L/(3*nelems^(1/3))")
    a = ap.parse_args()
    for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29544"), ("RANK", "0"), ("WORLD_SIZE", "1")):
        os.environ.setdefault(k, v)
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    R = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if R > 1 else "gloo",
                            device_id=torch.device("cuda", local) if R > 1 else None)
    P = importlib.import_module("pumi-pic_b200")
    comm = P.Comm()
    bx = 2 if R >= 2 else 1
    by = 2 if R >= 4 else 1
    bz = 2 if R >= 8 else 1
    n = a.cube_per_gpu
    # global cube: (bx*n) x (by*n) x (bz*n) would need a box mesh; use a cube of side n*cbrt-ish:
    # keep it a cube with N cells per edge and block owners by centroid
    N = int(round(n * (bx * by * bz) ** (1.0 / 3.0)))
    coords, ev = P.host_kuhn_cube(N, 1.0)
    e2s, s2v = P.host_derive_sides(3, ev)
    ne = ev.shape[0]
    cen = coords[ev].mean(axis=1)
    owner = ((cen[:, 0] * bx).astype(np.int64).clip(0, bx - 1)
             + bx * ((cen[:, 1] * by).astype(np.int64).clip(0, by - 1)
                     + by * (cen[:, 2] * bz).astype(np.int64).clip(0, bz - 1))).astype(np.int32) % R
    safe, part = P.host_picpart_tags(3, coords.shape[0], ev, owner, R, rank)
    gm = P.Mesh(3, coords, ev, e2s, s2v, np.ones(ne, np.int32))
    gm.set_picpart(safe, owner, rank)
    # particles: a.ppe per owned element, uniform in the tet, unit directions (generated on device)
    ppe = np.where(owner == rank, a.ppe, 0).astype(np.int32)
    members = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, members, ppe, elem_gids=np.arange(ne, dtype=np.int64))
    cap = ps.capacity
    lay = ps.layout()
    se = P.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps).long().clamp(0, ne - 1)
    g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
    w = -torch.log(torch.rand(cap, 4, device="cuda", dtype=torch.float64, generator=g).clamp_min(1e-12))
    w = w / w.sum(dim=1, keepdim=True)
    evd = torch.as_tensor(ev).cuda().long(); cod = torch.as_tensor(coords).cuda()
    pos = (cod[evd[se]] * w[:, :, None]).sum(dim=1)
    for k in range(3):
        ps.get(0)[k, :cap] = pos[:, k]
    d = torch.randn(cap, 3, device="cuda", dtype=torch.float64, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    for k in range(3):
        ps.get(3)[k, :cap] = d[:, k]
    ps.get(2)[0, :cap] = torch.arange(cap, dtype=torch.int32, device="cuda")
    del w, evd, cod, pos, d, se
    ext = 1.0
    push = a.push_mult * ext / (3 * ne ** (1.0 / 3))
    nverts = coords.shape[0]
    charge = torch.zeros(2 * nverts, dtype=torch.float64, device="cuda")
    names = ["push+search", "updatePtclPositions", "setUnsafeProcs", "migrate", "comm array reduce"]
    evs = {k: [] for k in names}
    sent_tot = 0
    step_evs = []

    def timed(name, fn, rec):
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record(); r = fn(); s1.record()
        if rec:
            evs[name].append((s0, s1))
        return r

    n0 = ps.nptcls
    for it in range(a.warmup + a.steps):
        rec = it >= a.warmup
        if it == a.warmup:
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            n_start = ps.nptcls
        b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
        b0.record()
        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        ids = torch.empty(max(ps.capacity, 1), dtype=torch.int32, device="cuda")
        sgn = push    # steady drift: particles stream across PICpart boundaries and out of the domain
        timed("push+search", lambda: P.push_direction_search(gm, ps, dr, sgn, x, tg, ids, elem_ids_empty=True,
                                                            from_orig=True, sync=False), rec)
        timed("updatePtclPositions", lambda: P.update_positions(ps, x, tg), rec)
        ne_d, np_d = timed("setUnsafeProcs", lambda: P.set_unsafe_procs(gm, ps, ids), rec)
        sent, recv = timed("migrate", lambda: P.migrate(ps, comm, ne_d, np_d), rec)
        timed("comm array reduce", lambda: comm.array_reduce(charge, nverts, 2, P.capi.PP_SUM), rec)
        b1.record()
        if rec:
            step_evs.append((b0, b1)); sent_tot += sent
    torch.cuda.synchronize()
    step_ms = float(np.sum([x.elapsed_time(y) for x, y in step_evs]))
    t = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(n_start), float(sent_tot), float(ps.nptcls)], dtype=torch.float64, device="cuda")
    if R > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    phases = {}
    for k in names:
        ms = torch.tensor([float(np.median([x.elapsed_time(y) for x, y in evs[k]]))], dtype=torch.float64,
                          device="cuda")
        if R > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        phases[k] = float(ms.item())
    if rank == 0:
        tot_ms = float(t.item())
        print(json.dumps({
            "bench": "full PIC step, weak scaling", "n_gpus": R, "steps": a.steps,
            "tets_global": int(ne), "tets_per_gpu": int(ne // R), "particles_global_start": cnt[0].item(),
            "particles_global_end": cnt[2].item(), "migrated_per_step": cnt[1].item() / a.steps,
            "ms_per_step_max_over_ranks": tot_ms / a.steps,
            "particle_steps_per_s": cnt[0].item() * a.steps / (tot_ms * 1e-3),
            "phase_median_ms_max_over_ranks": phases, "push_distance": push}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
