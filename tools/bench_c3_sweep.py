#!/usr/bin/env python
"""BASELINE configs[2]: the ps_combo160 rebuild / migrate sweep (performance_tests/ps_combo160.cpp).

For every (elements, particles) point of the two series of SURVEY 8(d) C3 --
  smallE_largeP: e = 1000 .. 5500 step 500,   p = 10 000 * e
  largeE_smallP: e = 10 000 .. 55 000 step 5000, p = 1000 * e
-- and every distribution of particle_structs/test/Distribute.cpp (1 uniform :77-89, 2 gaussian
mean ne/2 sigma ne/8 clamped :129-144, 3 exponential lambda 1 :171-215), a Sell-C-sigma structure of
160-byte particles (perfTypes.hpp:7-9; C = 32, sigma = ne, V = 1024) is rebuilt ITERS times, each
time with half of the particles re-drawn from the same distribution (ps_combo160.cpp:207-232); under
torchrun 10 % of the particles also go to a uniformly random other rank (migrate).  The random
streams are torch's, not Kokkos' XorShift pools (parity unpinned): the distributions are
the same, the draws are not.

    python tools/bench_c3_sweep.py [--quick] [--iters 20] [--series small,large] [--dists 1,2,3]
    torchrun --nproc-per-node N tools/bench_c3_sweep.py ...      (adds the migration)

Prints one JSON line per point: median ms per rebuild (or migrate), particles/s and GB/s at the
326 B per particle of SURVEY 8(d) (read + write of the 160-byte record, new_element, mask).
"""
import argparse
import ctypes as C
import importlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MEMBERS = [(np.float64, 17), (np.int32, 4), (np.int64, 1)]     # 136 + 16 + 8 = 160 bytes


def draw(torch, dist, ne, n, gen):
    """element per particle, int32 on the device"""
    if dist == 1:
        return torch.randint(0, ne, (n,), device="cuda", dtype=torch.int32, generator=gen)
    if dist == 2:
        e = torch.empty(n, device="cuda", dtype=torch.float32).normal_(ne / 2.0, ne / 8.0, generator=gen)
        return e.to(torch.int32).clamp_(0, ne - 1)               # int conversion truncates, then clamp
    # exponential (Distribute.cpp:171-215): uniform element -> inverse CDF, gaps filled uniformly
    lam = 1.0
    freq_max = -math.log(1.0 / ne)
    uni = torch.randint(0, ne, (n,), device="cuda", dtype=torch.int64, generator=gen)
    pct = uni.to(torch.float64) / ne
    start = (-1.0 / lam * torch.log(1 - pct) / freq_max * ne).to(torch.int64)
    nxt = torch.clamp(1 - pct - 1.0 / ne, min=1e-300)
    end = (-1.0 / lam * torch.log(nxt) / freq_max * ne).to(torch.int64)
    length = torch.clamp(end - start, min=1)
    inside = (torch.rand(n, device="cuda", generator=gen) * length).to(torch.int64)
    inside = torch.where(length > 1, torch.minimum(inside, length - 1), torch.zeros_like(inside))
    e = start + inside
    refill = torch.randint(0, ne, (n,), device="cuda", dtype=torch.int64, generator=gen)
    e = torch.where(e >= ne, refill, e)
    e = torch.where(uni == ne - 1, torch.zeros_like(e), e)
    return e.to(torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="first, middle and last point of each series")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--series", default="small,large")
    ap.add_argument("--dists", default="1,2,3")
    ap.add_argument("--percent-moved", type=float, default=0.5)
    ap.add_argument("--percent-moved-process", type=float, default=0.1)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    P = importlib.import_module("pumi-pic_b200")
    comm = P.Comm() if world > 1 else None
    if comm is not None:
        # 10 % of up to 55 M particles of 176 bytes (gid + record) may go to ONE peer in a step
        comm.set_p2p_window(int(0.13 * 55_000_000 * 176))
    series = {"small": [(e, 10000 * e) for e in range(1000, 5501, 500)],
              "large": [(e, 1000 * e) for e in range(10000, 55001, 5000)]}
    gen = torch.Generator(device="cuda")
    for sname in a.series.split(","):
        pts = series[sname]
        if a.quick:
            pts = [pts[0], pts[len(pts) // 2], pts[-1]]
        for ne, npt in pts:
            for d in [int(x) for x in a.dists.split(",")]:
                gen.manual_seed(1024 * 1024 + rank)                       # Distribute.h:27
                elems = draw(torch, d, ne, npt, gen)
                ppe = torch.bincount(elems.long(), minlength=ne).to(torch.int32).cpu().numpy()
                del elems
                gids = np.arange(ne, dtype=np.int64)
                ps = P.ParticleStructure(P.capi.PP_PS_SCS, MEMBERS, ppe, elem_gids=gids, sigma=ne, V=1024)
                times = []
                deferred = 0
                for it in range(a.iters + 1):
                    cap = ps.capacity
                    lay = ps.layout()
                    se = P.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps)
                    move = torch.rand(cap, device="cuda", generator=gen) < a.percent_moved
                    new = torch.where(move, draw(torch, d, ne, cap, gen), se)
                    procs = None
                    if world > 1:
                        go = torch.rand(cap, device="cuda", generator=gen) < a.percent_moved_process
                        other = torch.randint(0, world - 1, (cap,), device="cuda", dtype=torch.int32, generator=gen)
                        other = other + (other >= rank).to(torch.int32)     # a uniformly random OTHER rank
                        procs = torch.where(go, other, torch.full_like(other, rank))
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    t0.record()
                    if world > 1:
                        st = P.capi.MigrateStats()
                        P.capi.check(P.lib().pp_ps_migrate(ps.h, comm.h, P.api._ptr(new), P.api._ptr(procs), 0, None,
                                                           None, C.byref(st), P.api._stream()))
                        deferred += st.deferred
                    else:
                        ps.rebuild(new)
                    t1.record()
                    torch.cuda.synchronize()
                    if it:                                                  # first iteration = warm-up
                        times.append(t0.elapsed_time(t1))
                    del se, move, new, procs
                ms = float(np.median(times))
                n_now = ps.nptcls
                if world > 1:
                    tt = torch.tensor([ms, float(n_now)], dtype=torch.float64, device="cuda")
                    mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                    sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                    ms, n_now = float(mx[0]), int(sm[1])
                if rank == 0:
                    print(json.dumps({"config": "c3", "series": sname, "elements": ne, "particles_per_gpu": npt,
                                      "n_gpus": world, "distribution": {1: "uniform", 2: "gaussian", 3: "exponential"}[d],
                                      "op": "migrate" if world > 1 else "rebuild", "iters": a.iters,
                                      "ms_median": ms, "ms_min": float(min(times)),
                                      "particles_per_s": n_now / (ms * 1e-3),
                                      "GBps_at_326B": 326.0 * n_now / (ms * 1e-3) / 1e9,
                                      "max_ppe": int(ppe.max()), "deferred": int(deferred),
                                      "transport": ("peer-memory window" if comm.p2p_active else "NCCL") if comm else None}))
                    sys.stdout.flush()
                del ps
                torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
