#!/usr/bin/env python
"""Timing of the particle balancer's device phases on one B200 (CUDA events, median of 10):
addWeights (k_lb_count_ps), the plan (all-reduce skipped, host max-flow) and selectParticles
(k_lb_select_ps, two passes) for 10 M particles over 1 M elements acting as rank 0 of 8, every
element in one of 64 synthetic sbars.  Algorithmic bytes per slot: count 8.1 B (mask bit, new_proc,
new_elem; the element->vertex table is L2 resident), select 2 x 8.1 B + 4 B per re-targeted particle.
Prints one JSON line; copied into profiles/ by hand.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    P = importlib.import_module("pumi-pic_b200")
    R, ne, npt, nsb = 8, 1_000_000, 10_000_000, 64
    rng = np.random.default_rng(7)
    # sbar s: rank 0 plus a pseudo-random subset of the others; ids spaced by the number of parts
    table, gid = {}, 0
    for s in range(nsb):
        parts = (0,) + tuple(int(q) for q in range(1, R) if (s >> (q - 1)) & 1 or q == 1 + s % (R - 1))
        table[gid] = parts
        gid += len(parts)
    ids = np.asarray(sorted(table), np.int32)
    elem_sbar = ids[rng.integers(0, nsb, ne)]
    elem_owner = rng.integers(0, R, ne).astype(np.int32)
    bal = P.Balancer(R, 0, table, elem_sbar, elem_owner)
    ppe = np.full(ne, npt // ne, np.int32)
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, [(np.int32, 1)], ppe)
    cap = ps.capacity
    slot_elem, mask = ps.slot_elem_and_mask()
    new_elems = torch.as_tensor(np.where(mask.astype(bool), slot_elem, -1).astype(np.int32)).cuda()
    base_procs = torch.zeros(cap, dtype=torch.int32, device="cuda")

    def timed(fn, n=10):
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts[1:]))

    t_count = timed(lambda: bal.add_weights(ps, new_elems, base_procs))
    t0 = time.perf_counter()
    sends, imb = bal.balance(None, tol=1.05)
    t_plan = 1e3 * (time.perf_counter() - t0)
    sel = []
    for _ in range(6):
        bal.add_weights(ps, new_elems, base_procs)
        bal.balance(None, tol=1.05)
        procs = base_procs.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); bal.select(ps, new_elems, procs); b.record(); torch.cuda.synchronize()
        sel.append(a.elapsed_time(b))
    moved = int((procs != 0).sum().item())
    t_sel = float(np.median(sel[1:]))
    print(json.dumps({"workload": "balancer: %d particles, %d elements, %d sbars, rank 0 of %d" % (npt, ne, nsb, R),
                      "capacity": cap, "count_ms": t_count, "count_GBs": 8.125 * cap / t_count / 1e6,
                      "plan_ms_host": t_plan, "planned_sends": len(sends), "imbalance": imb,
                      "select_ms": t_sel, "select_GBs": (2 * 8.125 * cap + 4 * moved) / t_sel / 1e6,
                      "particles_retargeted": moved}))


if __name__ == "__main__":
    main()
