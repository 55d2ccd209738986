"""The full PIC step on N PICparts (BASELINE configs[4]), shared by bench.py and tools/bench_picstep.py.

Global mesh: a Kuhn cube whose elements are owned by bx*by*bz blocks (one PICpart core per GPU,
~6*cube^3 tets each).  PICparts follow the reference's defaults (pumipic_input.cpp:103-110): BFS
buffer with 3 layers, BFS safe zone with 1 layer.  With 2x2x2 blocks every core comes within three
layers of every other, whole cores are buffered (part_construct.cpp:407-437), so every rank holds
the full mesh and the ghost reduction is the all-reduce branch of Mesh::reduceCommArray
(pumipic_comm.cpp:234-247).

One step = the loop body of test/pseudoPushAndSearch.cpp:513-542 with the migration of
src/pumipic_ptcl_ops.hpp:73-85 and the field synchronisation of test/gyroScatter.hpp:231-258:
    fused push + search_mesh  ->  updatePtclPositions  ->  setUnsafeProcs
    ->  migrate (pp_ps_migrate over pp_comm: exchange + rebuild)  ->  comm-array all-reduce
Timed with CUDA events per phase on the current stream; the step time is the max over ranks.
"""
import importlib

import numpy as np

MEMBERS = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]
PHASES = ["push+search", "updatePtclPositions", "setUnsafeProcs", "migrate", "comm array reduce"]


def block_owner(coords, ev, R):
    """owner rank of every element: centroid -> block of a bx*by*bz grid"""
    bx = 2 if R >= 2 else 1
    by = 2 if R >= 4 else 1
    bz = 2 if R >= 8 else 1
    cen = coords[ev].mean(axis=1)
    own = ((cen[:, 0] * bx).astype(np.int64).clip(0, bx - 1)
           + bx * ((cen[:, 1] * by).astype(np.int64).clip(0, by - 1)
                   + by * (cen[:, 2] * bz).astype(np.int64).clip(0, bz - 1)))
    return (own % R).astype(np.int32)


class PicStep:
    def __init__(self, P, comm, rank, R, cube_per_gpu=55, ppe=10, push_mult=3.0, seed=1234, overlap_reduce=False, fuse_update=True):
        import torch
        self.P, self.comm, self.rank, self.R, self.torch = P, comm, rank, R, torch
        n = cube_per_gpu
        N = int(round(n * R ** (1.0 / 3.0)))
        coords, ev = P.host_kuhn_cube(N, 1.0)
        ne = ev.shape[0]
        owner = block_owner(coords, ev, R)
        if R > 1:
            # owner-major numbering, as a partitioned mesh has it (global ids are owner-major,
            # part_construct.cpp:335-374): every core is one contiguous range of elements, and the
            # vertices are numbered in the order the elements first touch them
            perm = np.argsort(owner, kind="stable")
            ev, owner = ev[perm], owner[perm]
            uniq, first = np.unique(ev.ravel(), return_index=True)
            new_id = np.empty(coords.shape[0], np.int32)
            new_id[uniq[np.argsort(first, kind="stable")]] = np.arange(uniq.shape[0], dtype=np.int32)
            ev = np.ascontiguousarray(new_id[ev])
            inv = np.empty_like(new_id)
            inv[new_id] = np.arange(new_id.shape[0], dtype=np.int32)
            coords = np.ascontiguousarray(coords[inv])
        e2s, s2v = P.host_derive_sides(3, ev)
        safe, part = P.host_picpart_tags(3, coords.shape[0], ev, owner, R, rank, P.api.BFS, P.api.BFS, 3, 1)
        assert part.all(), "block PICparts of this size buffer every core (full mesh on every rank)"
        self.gm = P.Mesh(3, coords, ev, e2s, s2v, np.ones(ne, np.int32))
        self.gm.set_picpart(safe, owner, rank)
        ppe_arr = np.where(owner == rank, ppe, 0).astype(np.int32)
        self.ps = ps = P.ParticleStructure(P.capi.PP_PS_SCS, MEMBERS, ppe_arr,
                                           elem_gids=np.arange(ne, dtype=np.int64))
        cap = ps.capacity
        lay = ps.layout()
        se = P.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps).long().clamp(0, ne - 1)
        g = torch.Generator(device="cuda")
        g.manual_seed(seed + rank)
        w = -torch.log(torch.rand(cap, 4, device="cuda", dtype=torch.float64, generator=g).clamp_min(1e-12))
        w = w / w.sum(dim=1, keepdim=True)
        evd = torch.as_tensor(ev).cuda().long()
        cod = torch.as_tensor(coords).cuda()
        pos = (cod[evd[se]] * w[:, :, None]).sum(dim=1)
        for k in range(3):
            ps.get(0)[k, :cap] = pos[:, k]
        d = torch.randn(cap, 3, device="cuda", dtype=torch.float64, generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        for k in range(3):
            ps.get(3)[k, :cap] = d[:, k]
        ps.get(2)[0, :cap] = torch.arange(cap, dtype=torch.int32, device="cuda")
        del w, evd, cod, pos, d, se
        self.push = push_mult * 1.0 / (3 * ne ** (1.0 / 3))
        self.nverts = coords.shape[0]
        self.ne = ne
        self.tets_per_gpu = ne // R
        self.charge = torch.zeros(2 * self.nverts, dtype=torch.float64, device="cuda")
        # fuse_update: updatePtclPositions (x <- xtgt, xtgt <- 0) is not a pass of its own but a member
        # remap of the record move that ends the migration (pp_ps_set_rebuild_remap)
        self.fuse_update = fuse_update
        # overlap_reduce (A/B option, off by default): the field synchronisation runs on its own stream;
        # nothing of the next step's push + search reads the reduced array (as in pseudoXGCm, whose push
        # does not use the field), so it may overlap with them; the next step's scatter point, and the
        # end of the run, wait for it.  Measured on 2 B200: no gain (1.86 vs 1.83 ms per step), the NCCL
        # kernel competes with the persistent walk kernel for SMs.
        self.overlap_reduce = overlap_reduce and R > 1
        self.comm_stream = torch.cuda.Stream() if self.overlap_reduce else None
        self.reduce_done = None

    def step(self, rec=None):
        """one PIC step; rec: dict phase -> list of (start, stop) events, or None"""
        P, ps, gm, torch = self.P, self.ps, self.gm, self.torch

        def timed(name, fn):
            if rec is None:
                return fn()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            rec[name].append((a, b))
            return r

        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        ids = torch.empty(max(ps.capacity, 1), dtype=torch.int32, device="cuda")
        timed("push+search", lambda: P.push_direction_search(gm, ps, dr, self.push, x, tg, ids, elem_ids_empty=True,
                                                            from_orig=True, sync=False))
        if self.fuse_update:
            timed("updatePtclPositions", lambda: ps.set_rebuild_remap([1, -1, 2, 3]))
        else:
            timed("updatePtclPositions", lambda: P.update_positions(ps, x, tg))
        ne_d, np_d = timed("setUnsafeProcs", lambda: P.set_unsafe_procs(gm, ps, ids))
        sent, recv = timed("migrate", lambda: P.migrate(ps, self.comm, ne_d, np_d))
        if not self.overlap_reduce:
            timed("comm array reduce", lambda: self.comm.array_reduce(self.charge, self.nverts, 2, P.capi.PP_SUM))
        else:
            main = torch.cuda.current_stream()
            if self.reduce_done is not None:         # a scatter of this step would write the array here:
                main.wait_event(self.reduce_done)    # the previous step's reduction must be through
            ready = torch.cuda.Event()
            ready.record(main)                       # the array is complete once the step's work is done
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                timed("comm array reduce", lambda: self.comm.array_reduce(self.charge, self.nverts, 2, P.capi.PP_SUM))
                self.reduce_done = torch.cuda.Event()
                self.reduce_done.record(self.comm_stream)
        return sent, recv

    # ---- full-size check of one step's record move (rebuild + migration), untimed
    _HK = (-7046029254386353131, -4417276706812531889, 1609587929392839161, -8796714831421723037,
           2685821657736338717, -3335678366873096957, 7046029254386353087, 6364136223846793005)

    def _record_hash(self, cols):
        """order-independent checksum of a set of records: every record (a tuple of int64 columns) is
        mixed into one 64-bit word (wrap-around arithmetic), the words are summed in two 31-bit halves
        so that neither a rank's sum nor the sum over ranks can overflow"""
        torch = self.torch
        acc = torch.zeros_like(cols[0])
        for k, c in enumerate(cols):
            c = c ^ (c >> 29)
            acc = (acc ^ (acc >> 31)) * 1099511628211 + c * self._HK[k % len(self._HK)]
        acc = acc ^ (acc >> 32)
        return torch.stack([(acc & 0x7fffffff).sum(), ((acc >> 31) & 0x7fffffff).sum()])

    def _mask_and_elements(self):
        torch, P, ps = self.torch, self.P, self.ps
        cap = ps.capacity
        lay = ps.layout()
        se = P.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps)
        mb = P.api._tensor_from_ptr(lay.mask_bits, ((cap + 31) // 32,), torch.int32, ps)
        sh = torch.arange(32, device="cuda", dtype=torch.int32)
        m = (((mb[:, None] >> sh[None, :]) & 1) != 0).reshape(-1)[:cap]
        return m, se

    def check_step(self, reduce_sum=None):
        """One more (untimed) step at FULL size with a complete check of what the rebuild and the migration
        did with the records: before the migration every surviving particle is the record (id, element the
        search found, target position, direction) -- bit patterns; after it every particle of the structure is
        (id, element of its row, position, direction).  The two multisets must be equal over all ranks
        (particles change rank, full-mesh PICparts number the elements alike everywhere): compared as the
        count and a 62-bit order-independent checksum.  Also: every target column is zero afterwards
        (updatePtclPositions).  reduce_sum: all-reduce (SUM) of an int64 cuda tensor over the ranks."""
        torch, P, ps, gm = self.torch, self.P, self.ps, self.gm
        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        cap = ps.capacity
        ids = torch.empty(max(cap, 1), dtype=torch.int32, device="cuda")
        P.push_direction_search(gm, ps, dr, self.push, x, tg, ids, elem_ids_empty=True, from_orig=True, sync=False)
        if self.fuse_update:
            ps.set_rebuild_remap([1, -1, 2, 3])
            newx = tg
        else:
            P.update_positions(ps, x, tg)
            newx = x
        ne_d, np_d = P.set_unsafe_procs(gm, ps, ids)
        m, _se = self._mask_and_elements()
        alive = m & (ne_d[:cap] >= 0)
        i64 = torch.int64

        def cols(pid, elem, pos, dirs, sel):
            c = [pid[sel].to(i64), elem[sel].to(i64)]
            c += [pos[k, :sel.shape[0]].contiguous().view(i64)[sel] for k in range(3)]
            c += [dirs[k, :sel.shape[0]].contiguous().view(i64)[sel] for k in range(3)]
            return c

        before = torch.cat([alive.sum().to(i64).reshape(1),
                            self._record_hash(cols(ps.get(2)[0, :cap], ne_d[:cap], newx, dr, alive))])
        P.migrate(ps, self.comm, ne_d, np_d)
        self.comm.array_reduce(self.charge, self.nverts, 2, P.capi.PP_SUM)
        cap2 = ps.capacity
        m2, se2 = self._mask_and_elements()
        x2, tg2, dr2 = ps.get(0), ps.get(1), ps.get(3)
        after = torch.cat([m2.sum().to(i64).reshape(1),
                           self._record_hash(cols(ps.get(2)[0, :cap2], se2, x2, dr2, m2))])
        tg_nonzero = sum(int((tg2[k, :cap2][m2] != 0).sum().item()) for k in range(3))
        local = torch.cat([before, after, torch.tensor([tg_nonzero, int(ps.nptcls)], dtype=i64, device="cuda")])
        if reduce_sum is not None:
            local = reduce_sum(local)
        v = [int(q) for q in local.cpu().tolist()]
        return {"particles_before": v[0], "particles_after": v[3], "count_mismatch": int(v[0] != v[3]),
                "record_checksum_mismatch": int(v[1:3] != v[4:6]), "targets_not_zeroed": v[6],
                "structure_count_mismatch": int(v[3] != v[7])}

    def finish(self):
        """the main stream waits for the last field synchronisation"""
        if self.reduce_done is not None:
            self.torch.cuda.current_stream().wait_event(self.reduce_done)

    def run(self, steps, warmup, barrier=None):
        """-> dict with per-rank timings (ms) of `steps` timed steps"""
        torch = self.torch
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        torch.cuda.synchronize()
        n_start = self.ps.nptcls
        rec = {k: [] for k in PHASES}
        sent_tot = 0
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            sent, _r = self.step(rec)
            sent_tot += sent
        self.finish()                                # includes the last (overlapped) field synchronisation
        t1.record()
        torch.cuda.synchronize()
        return {"step_ms_total": float(t0.elapsed_time(t1)),
                "phase_ms_median": {k: float(np.median([a.elapsed_time(b) for a, b in v])) for k, v in rec.items()},
                "particles_start": int(n_start), "particles_end": int(self.ps.nptcls), "sent": int(sent_tot)}


def run_picstep(P, comm, rank, R, steps, warmup, cube_per_gpu=55, ppe=10, push_mult=3.0, timing=False,
                overlap_reduce=False, fuse_update=True, full_size_check=2):
    """Collective over the ranks of torch.distributed (when R > 1).  Returns the record on rank 0.
    timing: also record the library's own phase timers (pp_timing_*, rank 0's table in the record)."""
    import torch
    import torch.distributed as dist
    ps = PicStep(P, comm, rank, R, cube_per_gpu, ppe, push_mult, overlap_reduce=overlap_reduce,
                 fuse_update=fuse_update)
    if timing:
        for _ in range(warmup):
            ps.step()
        warmup = 0
        torch.cuda.synchronize()
        P.api.timing_reset()
        P.api.timing_enable(True, rank)
    r = ps.run(steps, warmup, barrier=(dist.barrier if R > 1 else None))
    table = None
    if timing:
        table = {k: round(v["avg_ms"], 4) for k, v in P.api.timing_table().items()}
        P.api.timing_enable(False)
    t = torch.tensor([r["step_ms_total"]] + [r["phase_ms_median"][k] for k in PHASES], dtype=torch.float64,
                     device="cuda")
    cnt = torch.tensor([float(r["particles_start"]), float(r["sent"]), float(r["particles_end"])],
                       dtype=torch.float64, device="cuda")
    def allreduce(v, op):
        """torch.distributed plumbing; host tensors when the process group is gloo (ranks sharing one GPU)"""
        if dist.get_backend() == "nccl":
            dist.all_reduce(v, op=op)
            return v
        h = v.cpu()
        dist.all_reduce(h, op=op)
        return h.cuda()
    if R > 1:
        t = allreduce(t, dist.ReduceOp.MAX)
        cnt = allreduce(cnt, dist.ReduceOp.SUM)
    tot_ms = float(t[0].item())
    # algorithmic bytes of the full step per particle: push 49 + search 53 + mesh 7 +
    # position update 72 + setUnsafeProcs 4 + rebuild 110
    full_step_bytes = 295.0
    out = {"workload": "full PIC step (push, search, updatePtclPositions, setUnsafeProcs, migrate over pp_comm "
                       "[NCCL all-to-all-v + rebuild], comm-array all-reduce), block-partitioned Kuhn cube, "
                       "BFS 3-layer buffer / 1-layer safe zone, %d particles and %d tets per GPU"
                       % (ppe * ps.tets_per_gpu, ps.tets_per_gpu),
           "n_gpus": R, "steps": steps, "warmup": warmup, "tets_global": int(ps.ne),
           "particles_global_start": cnt[0].item(), "particles_global_end": cnt[2].item(),
           "migrated_per_step": cnt[1].item() / steps,
           "ms_per_step": tot_ms / steps,
           "value": cnt[0].item() * steps / (tot_ms * 1e-3), "unit": "particle full-steps/s",
           "phase_ms": {k: float(t[1 + i].item()) for i, k in enumerate(PHASES)},
           "algorithmic_bytes_per_particle_step": full_step_bytes,
           "achieved_GBs_per_gpu": full_step_bytes * cnt[0].item() / R / (tot_ms / steps * 1e-3) / 1e9,
           "scaling": "weak", "timing": "CUDA events around the K steps (incl. the last field synchronisation), max "
                                        "over ranks; phases: median over steps, max over ranks",
           "updatePtclPositions": ("folded into the record move of the migration's rebuild (pp_ps_set_rebuild_remap: "
                                   "x <- xtgt, xtgt <- 0 as a member remap; the phase time is the setter call)"
                                   if ps.fuse_update else "own pass (pp_update_positions)"),
           "comm_array_reduce": ("on its own stream, overlapped with the next step's push + search + migration"
                                 if ps.overlap_reduce else "in line"),
           "transport": "peer-memory window (NVLink P2P stores, no host round trip)" if comm.p2p_active
                        else ("NCCL AllGather + grouped Send/Recv" if R > 1 else "single rank")}
    if table is not None:
        out["library_phase_avg_ms_rank0"] = table
    if full_size_check:
        chk = [ps.check_step((lambda v: allreduce(v, dist.ReduceOp.SUM)) if R > 1 else None)
               for _ in range(full_size_check)]
        out["full_size_check"] = {
            "steps": len(chk), "particles_per_step": [c["particles_after"] for c in chk],
            "count_mismatch": sum(c["count_mismatch"] + c["structure_count_mismatch"] for c in chk),
            "record_checksum_mismatch": sum(c["record_checksum_mismatch"] for c in chk),
            "targets_not_zeroed": sum(c["targets_not_zeroed"] for c in chk),
            "what": "untimed extra steps at full size: the multiset of records (id, element, position, direction: bit "
                    "patterns) of the surviving particles before the migration equals the multiset in the rebuilt "
                    "structures after it, over all ranks (count + 62-bit order-independent checksum); element = the "
                    "row's element after, the search's result before; targets are zero afterwards"}
    del ps
    torch.cuda.empty_cache()
    return out
