#!/usr/bin/env python
"""bench.py -- particle push+search throughput of the B200-native PUMI-PIC hot path.

Workload (BASELINE.json configs[1], "search_mesh_3d adjacency walk only, 10M particles on
1M-tet synthetic cube mesh, 1xB200"): Kuhn cube N=55 (998 250 tets), 10 M particles placed by
test_adj.cpp's seeded generator, push distance L/(3*nelems^(1/3)).  One step = ONE fused kernel:
xtgt = x + d*dir followed by the search_mesh BCC walk; the two position buffers swap roles and
the sign of d flips every step so particles oscillate and the population is stationary after
the first step (particles that leave the domain are deleted, as in the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

CPU legs (`cpu_baseline`, `--impl reference`): the reference's own push_ptcls + search_mesh source,
compiled unmodified into oracle/_ref over OpenMP stand-ins for Kokkos / Omega_h (kind "reference");
the OpenMP oracle port where that library is missing (kind "port").

N>1 (torchrun): every rank owns an identical-size independent shard (its own PICpart-sized
mesh + particles); push+search has no exchange step, so there is no data-path collective and
scaling is weak.  Prints ONE JSON line on rank 0.

The same line carries `picstep`: the FULL PIC step of BASELINE configs[4] (push, search,
updatePtclPositions, setUnsafeProcs, migrate over the library's own NCCL communicator, comm-array
all-reduce; pumi-pic_b200/picstep.py) on the same N GPUs, with per-phase times, the particles
migrated per step, an in-run parity check of the multi-rank loop against the serial oracle (small
case) and `full_size_check` (two untimed steps at full size: the multiset of particle records before
the migration equals the multiset in the rebuilt structures after it, count + checksum); and
`parity`: the element ids of one step of the headline workload at full size compared with the
reference's own search_mesh source on the same inputs.
"""
import argparse
import importlib
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle push+search steps/s"
UNIT = "particle-steps/s"
# push 49 B + BCC search 53 B + mesh 7 B per particle-step (unfused accounting)
ALGO_BYTES_PER_PARTICLE_STEP = 109.0
MEMBERS = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class HostMesh:
    pass


def build_workload(pp, wl, cube_n, nptcls):
    """Host-side synthetic inputs (seeded, identical on every rank and for the CPU arm)."""
    coords, ev = pp.host_kuhn_cube(cube_n, 1.0)
    e2s, s2v = pp.host_derive_sides(3, ev)
    m = HostMesh()
    m.dim, m.coords, m.elem2verts, m.elem2sides, m.side2verts = 3, coords, ev, e2s, s2v
    m.nelems, m.nverts, m.nsides = ev.shape[0], coords.shape[0], s2v.shape[0]
    m.class_id = np.ones(m.nelems, np.int32)
    ppe = wl.even_ppe(m.nelems, nptcls)
    return m, ppe


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Polls NVML (SM clock, max clock, power, clock-event reasons) from a thread while the timed
    region runs; the steps take milliseconds, far below nvidia-smi's start-up time."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        import threading
        self.samples = []
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.window = [None, None]
        self.t.start()

    def _run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                break
            time.sleep(0.001)

    def begin(self):
        self.window[0] = time.perf_counter()

    def end(self):
        self.window[1] = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.h is None or not self.samples:
            return out
        lo, hi = self.window
        inside = [s for s in self.samples if lo is not None and lo <= s[0] <= hi]
        use = inside if inside else self.samples[-5:]
        reasons = set()
        for _, _, rs, _ in use:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        out.update(sm_mhz=float(np.median([s[1] for s in use])), sm_max_mhz=float(self.max_sm),
                   reasons=sorted(reasons), samples=len(use),
                   power_w_max=float(max(s[3] for s in use)),
                   window="timed region" if inside else "last samples before the end of the timed region")
        return out


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpumipic_ref_primitives.so")
REF_NOTE = ("the reference's own push_ptcls (test/test_adj.cpp:550-562) + search_mesh (src/pumipic_adjacency.tpp:642) "
            "source, compiled unmodified into oracle/_ref over OpenMP stand-ins for Kokkos / Omega_h "
            "(oracle/ref_shim); Omega_h's per-call mesh derivations are precomputed, which favours the reference")
PORT_NOTE = "OpenMP oracle (CPU restatement of the reference algorithm)"


def load_ref_lib():
    """oracle/_ref: the reference's search source compiled from /root/reference in the authoring
    container (oracle/build_ref_primitives.py); the prebuilt library travels to the GPU box."""
    import ctypes as C
    if not os.path.exists(REF_LIB):
        return None
    try:
        L = C.CDLL(REF_LIB)
        L.ref_bench_create.restype = C.c_void_p
        L.ref_get_max_threads.restype = C.c_int
    except (OSError, AttributeError):          # built elsewhere and not loadable here: use the port
        return None
    return L


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_leg(orc, om, wl, m, ppe, sample, steps, warmup, ref=None, loop="seeded", given=None,
                      threads=None):
    """The reference's CPU path: push (xtgt = x + d*dir) + search_mesh BCC per step, in the same loop
    form as the GPU arm (seeded: every step starts from the same positions with elem_ids seeded from
    the structure rows, +d / -d alternating; pingpong: buffers swap, ids carry over).  With `ref`
    (oracle/_ref) the reference's own source runs; without it the OpenMP oracle port.
    given = (slot_elem, mask, X, D): run on exactly these slot arrays (the GPU arm's structure)
    instead of a dense array of the first `sample` particles.  Returns (particle-steps/s, ms/step,
    slots, ids after the first timed-or-not step with +d)."""
    import ctypes as C
    threads = threads or host_threads()
    # torchrun exports OMP_NUM_THREADS=1: set the team size explicitly, not from the environment
    if ref is not None:
        ref.ref_set_num_threads(int(threads))
    orc.lib().orc_set_num_threads(int(threads))
    if given is not None:
        slot_elem, mask, X, D = given
        slot_elem = np.ascontiguousarray(slot_elem, np.int32)
        mask = np.ascontiguousarray(mask, np.uint8)
        cap = int(mask.shape[0])
    else:
        cap = int(sample)
        slot_elem = np.repeat(np.arange(m.nelems, dtype=np.int32), ppe)[:cap]
        mask = np.ones(cap, np.uint8)
        X, D = wl.init3d_internal(m, slot_elem, mask)
    dist = wl.push_distance(m)
    A, B = X, np.zeros_like(X)
    ids = None
    handle = None
    first_ids = None
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    if ref is not None:
        off, val = om.side2elem_off(), om.side2elem()
        exposed = np.ascontiguousarray(om.exposed(), np.int8)
        handle = C.c_void_p(ref.ref_bench_create(
            3, m.nverts, m.coords.ctypes.data_as(dp), m.nelems, m.elem2verts.ctypes.data_as(ip), m.nsides,
            m.elem2sides.ctypes.data_as(ip), m.side2verts.ctypes.data_as(ip), off.ctypes.data_as(ip),
            val.ctypes.data_as(ip), exposed.ctypes.data_as(C.POINTER(C.c_byte)), om.vol().ctypes.data_as(dp),
            cap, slot_elem.ctypes.data_as(ip), mask.ctypes.data_as(C.POINTER(C.c_ubyte))))
        ids = np.full(cap, -1, np.int32)
    seeded = loop == "seeded"
    times, active = [], []
    for it in range(warmup + steps):
        sgn = dist if it % 2 == 0 else -dist
        fresh = seeded or it == 0
        t0 = time.perf_counter()
        if handle is not None:
            ref.ref_bench_step(handle, A.ctypes.data_as(dp), B.ctypes.data_as(dp), D.ctypes.data_as(dp),
                               C.c_long(A.shape[1]), C.c_double(sgn), ids.ctypes.data_as(ip), int(fresh))
        else:
            np.copyto(B, A)                       # xtgt = x ...
            orc.push_direction(mask, B, D, sgn)   # ... + d*dir   (same arithmetic as the fused push)
            found, ids, _, _, st = om.search_mesh(slot_elem, mask, A, B, elem_ids=None if fresh else ids)
        dt = time.perf_counter() - t0
        if it == 0:
            first_ids = ids.copy()
        if it >= warmup:
            times.append(dt)
            active.append(int(mask.sum()) if seeded else int((ids >= 0).sum()))
        if not seeded:
            A, B = B, A
    if handle is not None:
        ref.ref_bench_destroy(handle)
    return float(sum(active)) / sum(times), 1e3 * sum(times) / len(times), cap, first_ids


def numpy_workload(wl, cube_n, nptcls):
    """The workload's mesh built without the product library (tests/meshes.py, numpy): the
    reference arm must not load libpumipic_b200.so.  Identical arrays (tests/test_capi_loads.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from meshes import kuhn_cube
    tm = kuhn_cube(cube_n)
    return tm, wl.even_ppe(tm.nelems, nptcls)


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads (and so its first-touch pinned buffers) to the CPUs NVML reports
    as local to its GPU; without it all ranks of a box stage through one memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def scale_parity(ids_gpu, ids_ref, mask, om, Xnp, Dnp, dist, tol=1e-9):
    """Element ids of the GPU step vs the reference's on the same slots.  near_face: mismatching
    particles whose target lies within `tol` (barycentric) of a face of the reference's element --
    none are expected, the geometry is bit-identical (-fmad=false)."""
    m = mask.astype(bool)
    diff = np.nonzero(m & (ids_gpu != ids_ref))[0]
    near = 0
    if diff.size:
        for s_ in diff[:1000]:
            e = int(ids_ref[s_])
            if e < 0:
                continue
            v = om.mesh.coords[om.mesh.elem2verts[e]]
            tgt = Xnp[:, s_] + dist * Dnp[:, s_]
            lam = np.linalg.solve((v[1:] - v[0]).T, tgt - v[0])
            b = np.concatenate([[1.0 - lam.sum()], lam])
            if np.min(np.abs(b)) < tol:
                near += 1
    return {"compared": int(m.sum()), "mismatch": int(diff.size), "near_face": int(near), "epsilon": tol,
            "against": "reference search_mesh source (oracle/_ref) on the same slots, first +d step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cube-n", type=int, default=55)
    ap.add_argument("--particles", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles per CPU step (0 = all)")
    ap.add_argument("--ps", default="scs", choices=["dps", "scs", "csr"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-picstep", action="store_true")
    ap.add_argument("--picstep-steps", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the timed steps eagerly (one event per step) instead of one CUDA graph")
    ap.add_argument("--e2e-parts", type=int, default=8, help="pieces of the pipelined host-buffer step")
    ap.add_argument("--walk-kernel", type=int, default=2, choices=[0, 1, 2],
                    help="0 thread-per-slot, 1 block-staged, 2 Sell-C-sigma chunk walk (default)")
    ap.add_argument("--loop", default="seeded", choices=["seeded", "pingpong"],
                    help="seeded: every step pushes from the rebuilt positions with elem_ids seeded "
                         "from the structure rows (what every reference call site does); "
                         "pingpong: buffers swap and elem_ids carry over, no rebuild in between")
    a = ap.parse_args()
    if a.cpu_sample <= 0:
        a.cpu_sample = a.particles
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    wl = load_module("pp_workloads", os.path.join(ROOT, "pumi-pic_b200", "workloads.py"))
    config = {"workload": "search_mesh BCC walk + push, %d particles on Kuhn cube N=%d (%d tets) per GPU"
                          % (a.particles, a.cube_n, 6 * a.cube_n ** 3),
              "particles_per_gpu": a.particles, "tets_per_gpu": 6 * a.cube_n ** 3,
              "push": "xtgt = x + d*dir, d = L/(3*nelems^(1/3)), sign alternates per step",
              "l2": "inputs (>=0.8 GB of particle columns per step) are larger than the 126 MB L2",
              "particle_structure": a.ps, "loop": a.loop,
              "note": "the timed loop is the fused push+search on a freshly rebuilt structure (ids seeded from the "
                      "rows, positions not advanced); the step WITH position update, rebuild and migration is "
                      "the `picstep` record of the same line",
              "launch": "eager, one CUDA event per step" if a.no_graph else "the K timed steps replayed as one CUDA graph",
              "parallelism": "independent shard per GPU (no exchange in push+search)"}

    if a.impl == "reference":
        # CPU arm: rank 0 only; the reference's own push_ptcls + search_mesh source (oracle/_ref), or
        # the oracle port where that library is missing.  Nothing of the product library is loaded:
        # the mesh comes from tests/meshes.py (numpy).
        if rank != 0:
            return 0
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api as orc
        tm, ppe = numpy_workload(wl, a.cube_n, a.particles)
        om = orc.OracleMesh(tm)
        ref = load_ref_lib()
        cores = host_threads()
        val, ms, cap, _ = cpu_reference_leg(orc, om, wl, tm, ppe, a.cpu_sample, a.steps, a.warmup, ref=ref,
                                            loop=a.loop, threads=cores)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores,
                                 "kind": "reference" if ref is not None else "port",
                                 "sample": "first %d particles of the workload per step, %s loop; %s"
                                           % (cap, a.loop, REF_NOTE if ref is not None else PORT_NOTE)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pp = importlib.import_module("pumi-pic_b200")
    P = pp
    P.lib().pp_search_set_staged(a.walk_kernel)
    m, ppe = build_workload(pp, wl, a.cube_n, a.particles)
    gm = pp.Mesh(3, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
    kind = {"dps": P.capi.PP_PS_DPS, "scs": P.capi.PP_PS_SCS, "csr": P.capi.PP_PS_CSR}[a.ps]
    ps = pp.ParticleStructure(kind, MEMBERS, ppe)
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = wl.init3d_internal(m, slot_elem, mask)
    d = wl.push_distance(m)
    cap = ps.capacity
    xa = torch.as_tensor(X).cuda()
    xb = torch.zeros_like(xa)
    dr = torch.as_tensor(D).cuda()
    ids = torch.zeros(cap, dtype=torch.int32, device="cuda")

    seeded = a.loop == "seeded"

    def step(it, a_, b_, sync=False):
        if seeded:   # x stays in xa (the rebuilt state); +d / -d alternate so no step repeats its predecessor
            return P.push_direction_search(gm, ps, dr, d if it % 2 == 0 else -d, xa, xb, ids,
                                           elem_ids_empty=True, from_orig=True, sync=sync)
        return P.push_direction_search(gm, ps, dr, d if it % 2 == 0 else -d, a_, b_, ids,
                                       elem_ids_empty=False, from_orig=True, sync=sync)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # the first +d step from the initial positions: its element ids are what `parity` compares with
    # the reference arm's first step on the same slots
    P.push_direction_search(gm, ps, dr, d, xa, xb, ids, elem_ids_empty=True, from_orig=True, sync=True)
    ids_first = ids.cpu().numpy().copy() if (rank == 0 and world == 1 and not a.no_cpu_baseline) else None
    # seed element ids from the structure rows, then warm up
    P.push_direction_search(gm, ps, dr, 0.0, xa, xb, ids, elem_ids_empty=True, from_orig=True, sync=True)
    A, B = xa, xb
    it = 0
    for _ in range(a.warmup):
        step(it, A, B); A, B = B, A; it += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The K timed steps are captured once in a CUDA graph (K fused-kernel launches with their
    # alternating push sign) and replayed by ONE launch: the step is 0.27 ms of GPU work, and with
    # N ranks sharing the host the Python/ctypes launch path otherwise shows up as gaps between
    # kernels.  --no-graph times the same K launches eagerly with an event after every step.
    stream = torch.cuda.current_stream()
    graph = None
    if not a.no_graph:
        cs = torch.cuda.Stream()
        cs.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cs):
            for k in range(a.steps):
                step(it, A, B); A, B = B, A; it += 1
        graph.replay()            # untimed: uploads the graph, extra warm-up of the same K steps
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    if sampler:
        sampler.begin()
    if graph is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms = e0.elapsed_time(e1)
        kernel_ms = [total_ms / a.steps] * a.steps
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record(stream)
        for k in range(a.steps):
            step(it, A, B); A, B = B, A; it += 1
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
        total_ms = ev[0].elapsed_time(ev[-1])
        kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    if sampler:
        sampler.end()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    st = P.capi.SearchStats()
    P.capi.check(P.lib().pp_search_last_stats(gm.h, st, None))
    # particles pushed + searched per step: every masked particle in the seeded loop; in the
    # ping-pong loop the ones still inside the domain (the oscillation keeps that set constant)
    live = int(st.active) if seeded else int((ids >= 0).sum().item())
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    n = torch.tensor([float(live)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    live_all = float(n.item())
    value = live_all * a.steps / (total_ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST (pinned) buffers: H2D of the step's inputs and
    # D2H of its results inside the timed region.  The direction column is an input of the
    # structure, not of the step: it is uploaded by the first (untimed) call and stays resident
    # (h_dir = NULL afterwards), as a caller that keeps `dir` on the device between steps would.
    e2e = None
    if not a.no_e2e:
        hx = torch.empty_like(A, device="cpu").pin_memory(); hx.copy_(A)
        hd = torch.empty_like(dr, device="cpu").pin_memory(); hd.copy_(dr)
        hi = torch.empty_like(ids, device="cpu").pin_memory(); hi.copy_(ids)
        ht = torch.empty_like(A, device="cpu").pin_memory()
        dx, dd, dt_, di = (torch.empty_like(A), torch.empty_like(dr), torch.empty_like(A),
                           torch.empty_like(ids))
        e2e_steps = max(3, min(a.steps, 5))

        def e2e_step(k):
            if seeded:
                # the reference-facing call with HOST buffers: copies and kernel pipelined inside
                P.push_direction_search_host(gm, ps, hx, hd if k == 0 else None, ht, hi,
                                             d if (it + k) % 2 == 0 else -d, nparts=a.e2e_parts, sync=False)
                torch.cuda.synchronize()
                return
            dx.copy_(hx, non_blocking=True); dd.copy_(hd, non_blocking=True)
            di.copy_(hi, non_blocking=True)
            P.push_direction_search(gm, ps, dd, d if (it + k) % 2 == 0 else -d, dx, dt_, di,
                                    elem_ids_empty=False, from_orig=True, sync=False)
            ht.copy_(dt_, non_blocking=True); hi.copy_(di, non_blocking=True)
            torch.cuda.synchronize()
            hx.copy_(ht)   # the caller's next step starts from the pushed positions
        e2e_step(0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(1, e2e_steps + 1):
            e2e_step(k)
        torch.cuda.synchronize()
        et = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        live2 = torch.tensor([float(live if seeded else (hi >= 0).sum().item())], dtype=torch.float64,
                             device="cuda")
        if world > 1:
            dist.all_reduce(live2, op=dist.ReduceOp.SUM)
        e2e = {"value": float(live2.item()) * e2e_steps / float(et.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(hx.numel() * 8 + (0 if seeded else hd.numel() * 8 + hi.numel() * 4)),
               "d2h_bytes_per_step": int(ht.numel() * 8 + hi.numel() * 4),
               "steps": e2e_steps, "host_cpus_bound": numa_cpus,
               "note": ("pp_push_direction_search_host: pinned host buffers, %d pieces, H2D / kernel / D2H "
                        "overlapped on three streams; positions in, targets + element ids out every step; the "
                        "direction column is uploaded once (untimed first call) and stays resident"
                        % a.e2e_parts) if seeded else
                       "pinned host buffers, cudaMemcpyAsync H2D -> fused kernel -> D2H per step"}
        del hx, hd, hi, ht, dx, dd, dt_, di

    # ---- the full PIC step on the same GPUs (BASELINE configs[4]) + parity of the multi-rank loop
    picstep = None
    if not a.no_picstep:
        del xa, xb, dr, ids, A, B
        graph = None
        del ps, gm
        torch.cuda.empty_cache()
        ps_mod = importlib.import_module("pumi-pic_b200.picstep")
        comm = P.Comm()
        picstep = ps_mod.run_picstep(P, comm, rank, world, a.picstep_steps, 3, cube_per_gpu=a.cube_n,
                                     ppe=max(1, a.particles // (6 * a.cube_n ** 3)))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        try:
            import mgpu_worker
            par = mgpu_worker.pic_loop_parity(P, comm, rank, world, steps=4, fuse_update=True)
        except Exception as ex:  # the checker must not take the measurement down with it
            par = {"error": repr(ex)[:200]}
        if picstep is not None:
            picstep["parity"] = par
        del comm

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    kavg_ms = float(np.mean(kernel_ms))
    live_rank0 = live
    achieved = ALGO_BYTES_PER_PARTICLE_STEP * live_rank0 / (kavg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("k_search_dram_bytes_per_launch")
        traffic_src = tj.get("source")
    cpu, parity = None, None
    if not a.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_api as orc
        from meshes import Mesh as TMesh
        tm = TMesh(3, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
        om = orc.OracleMesh(tm)
        ref = load_ref_lib()
        cores = host_threads()
        # the same slots, positions and directions as the GPU arm (the whole structure), same loop
        val, ms, ncap, ids_ref = cpu_reference_leg(orc, om, wl, tm, ppe, cap, 3, 1, ref=ref, loop=a.loop,
                                                   given=(slot_elem, mask, X, D), threads=cores)
        cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": "reference" if ref is not None else "port",
               "sample": "the GPU arm's own structure (%d slots, %d particles), %s loop, 3 timed steps after 1 "
                         "warm-up; %s" % (ncap, int(mask.sum()), a.loop, REF_NOTE if ref is not None else PORT_NOTE),
               "ms_per_step": ms}
        if ids_first is not None and ids_ref is not None:
            parity = scale_parity(ids_first, ids_ref, mask, om, X, D, d)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": a.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": ["k_search<3,BCC,PUSH>", "k_walk_bcc<3,PUSH>", "k_walk_scs<3,PUSH>"][a.walk_kernel]
                                   + " (fused push + walk)",
                         "algorithmic_bytes_per_particle_step": ALGO_BYTES_PER_PARTICLE_STEP,
                         "kernel_ms": kavg_ms, "peak_source": peak_src},
            "cpu_baseline": cpu, "parity": parity, "picstep": picstep,
            "detail": {"live_particles_per_gpu": live_rank0, "capacity": cap,
                       "walk_iterations_last_step": st.loops, "hops_last_step": int(st.hops),
                       "active_last_step": st.active}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
