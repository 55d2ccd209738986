#!/usr/bin/env python
"""Phase breakdown of the full PIC step on one B200 (CUDA events per phase, median over steps).

  c2 : Kuhn cube N=55 (998 250 tets), 10 M particles, Sell-C-sigma: fused push+search,
       updatePtclPositions, rebuild  (BASELINE configs[1] as a real time-step loop)
  c3 : ps_combo160-like rebuild sweep: 160-byte particles, 50 % of the particles re-drawn per step
  c4 : 2D plate (2 M triangles), 50 M particles: push + search_mesh_2d + rebuild + 2 x gyroScatter

Prints one JSON object per configuration; results are copied into profiles/ by hand.
"""
import argparse
import importlib
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class HostMesh:
    pass


def host_mesh(pp, dim, n):
    coords, ev = (pp.host_kuhn_cube(n, 1.0) if dim == 3 else pp.host_plate(n, 1.0))
    e2s, s2v = pp.host_derive_sides(dim, ev)
    m = HostMesh()
    m.dim, m.coords, m.elem2verts, m.elem2sides, m.side2verts = dim, coords, ev, e2s, s2v
    m.nelems, m.nverts = ev.shape[0], coords.shape[0]
    m.class_id = np.ones(m.nelems, np.int32)
    return m


class Timer:
    def __init__(self, torch):
        self.torch = torch
        self.t = {}

    def run(self, name, fn):
        a = self.torch.cuda.Event(enable_timing=True)
        b = self.torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.t.setdefault(name, []).append((a, b))
        return r

    def summary(self, skip=1):
        self.torch.cuda.synchronize()
        out = {}
        for k, evs in self.t.items():
            ms = [a.elapsed_time(b) for a, b in evs]
            ms = ms[skip:] if len(ms) > skip else ms
            out[k] = {"median_ms": float(np.median(ms)), "min_ms": float(np.min(ms)), "n": len(ms)}
        return out


def c2(pp, wl, torch, steps, nptcls, cube_n, kind):
    P = pp
    m = host_mesh(pp, 3, cube_n)
    gm = pp.Mesh(3, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
    members = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]
    ps = pp.ParticleStructure(kind, members, wl.even_ppe(m.nelems, nptcls))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = wl.init3d_internal(m, slot_elem, mask)
    cap = ps.capacity
    ps.get(0)[:, :cap] = torch.as_tensor(X).cuda()
    ps.get(3)[:, :cap] = torch.as_tensor(D).cuda()
    ps.get(2)[0, :cap] = torch.arange(cap, dtype=torch.int32, device="cuda")
    d = wl.push_distance(m)
    T = Timer(torch)
    counts = []
    for it in range(steps + 1):
        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        ids = torch.empty(ps.capacity, dtype=torch.int32, device="cuda")
        T.run("push+search (fused)", lambda: P.push_direction_search(gm, ps, dr, d, x, tg, ids,
                                                                     elem_ids_empty=True, from_orig=True,
                                                                     sync=False))
        T.run("updatePtclPositions", lambda: P.update_positions(ps, x, tg))
        T.run("rebuild", lambda: ps.rebuild(ids))
        counts.append(ps.nptcls)
    s = T.summary()
    tot = sum(v["median_ms"] for v in s.values())
    return {"config": "c2", "particles": nptcls, "tets": m.nelems, "phases": s,
            "full_step_ms": tot, "full_steps_per_s": nptcls / (tot * 1e-3),
            "particles_after": counts[-1], "capacity": ps.capacity}


def c3(pp, wl, torch, steps, ne, np_, kind):
    """ps_combo160.cpp:207-232: every step half of the particles get a new random element."""
    members = [(np.float64, 17), (np.int32, 4), (np.int64, 1)]       # perfTypes.hpp:7-9, 160 B
    rng = np.random.default_rng(1024 * 1024)
    ppe = np.bincount(rng.integers(0, ne, np_), minlength=ne).astype(np.int32)
    ps = pp.ParticleStructure(kind, members, ppe, sigma=ne, V=1024)
    T = Timer(torch)
    for it in range(steps + 1):
        cap = ps.capacity
        lay_se, _ = None, None
        lay = ps.layout()
        se = pp.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps)
        move = torch.rand(cap, device="cuda") < 0.5
        new = torch.where(move, torch.randint(0, ne, (cap,), device="cuda", dtype=torch.int32), se)
        T.run("rebuild", lambda: ps.rebuild(new))
    s = T.summary()
    return {"config": "c3", "elements": ne, "particles": np_, "record_bytes": 160, "phases": s,
            "rebuild_particles_per_s": np_ / (s["rebuild"]["median_ms"] * 1e-3),
            "rebuild_GBps_at_326B": 326.0 * np_ / (s["rebuild"]["median_ms"] * 1e-3) / 1e9}


def c4(pp, wl, torch, steps, nptcls, plate_n, kind):
    P = pp
    m = host_mesh(pp, 2, plate_n)
    gm = pp.Mesh(2, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
    members = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]
    # the reference's own load (pseudoXGCm.cpp:167-222, pinned against its compiled source): a normal
    # number of particles per triangle, the rounding shortfall in the last one
    ppe, _ = wl.xgc_source_elements(m.class_id, np.zeros(m.nelems, np.int32), 0, 141, nptcls)
    ps = pp.ParticleStructure(kind, members, ppe)
    cap = ps.capacity
    lay = ps.layout()
    se = pp.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps).long().clamp(0, m.nelems - 1)
    # positions: element centroid + small in-element jitter, directions uniform on the circle (device side)
    ev = torch.as_tensor(m.elem2verts).cuda().long()
    co = torch.as_tensor(m.coords).cuda()
    w = torch.rand(cap, 3, device="cuda", dtype=torch.float64) + 0.05
    w = w / w.sum(dim=1, keepdim=True)
    pos = (co[ev[se]] * w[:, :, None]).sum(dim=1)                  # [cap, 2]
    ps.get(0)[0, :cap] = pos[:, 0]; ps.get(0)[1, :cap] = pos[:, 1]; ps.get(0)[2, :cap] = 0
    ang = torch.rand(cap, device="cuda", dtype=torch.float64) * 2 * np.pi
    ps.get(3)[0, :cap] = torch.cos(ang); ps.get(3)[1, :cap] = torch.sin(ang); ps.get(3)[2, :cap] = 0
    del ev, co, w, pos, ang, se
    d = wl.push_distance(m)
    rings, ppr, rmax = 3, 8, 0.038 * (1.0 / plate_n) / 0.01       # a few element widths
    T = Timer(torch)
    fmap, st = T.run("gyro ring map (setup)", lambda: P.gyro_ring_map(gm, rmax, rings, ppr, 0.0))
    for it in range(steps + 1):
        x, tg, dr = ps.get(0), ps.get(1), ps.get(3)
        ids = torch.full((ps.capacity,), -1, dtype=torch.int32, device="cuda")
        T.run("push", lambda: P.push_from(ps, x, tg, dr, d))
        T.run("search_mesh_2d", lambda: P.search_mesh(gm, ps, x, tg, ids, variant=P.capi.PP_SEARCH_2D_LEGACY,
                                                       elem_ids_empty=True, looplimit=200, sync=False))
        T.run("updatePtclPositions", lambda: P.update_positions(ps, x, tg))
        T.run("rebuild", lambda: ps.rebuild(ids))
        T.run("gyroScatter x2", lambda: (P.gyro_scatter(gm, ps, fmap, rmax, rings, ppr),
                                         P.gyro_scatter(gm, ps, fmap, rmax, rings, ppr)))
    s = T.summary()
    tot = sum(v["median_ms"] for k, v in s.items() if "setup" not in k)
    return {"config": "c4", "particles": nptcls, "triangles": m.nelems, "verts": m.nverts, "phases": s,
            "full_step_ms": tot, "full_steps_per_s": nptcls / (tot * 1e-3), "particles_after": ps.nptcls}


def scatter_conservation(pp, torch, gm, ps, m, fmap, rmax, rings, ppr):
    """full-size property of the gyro scatter (gyroScatter.hpp:168-229): every particle puts 1 on rings 0 and 1
    of its element's 3 vertices, and a ring's value reaches the mesh once per mapped (point, vertex) pair divided
    by the points per ring -- the total over the vertices is known from the map and the particles per element"""
    cap = ps.capacity
    lay = ps.layout()
    se = pp.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps).long()
    mb = pp.api._tensor_from_ptr(lay.mask_bits, ((cap + 31) // 32,), torch.int32, ps)
    sh = torch.arange(32, device="cuda", dtype=torch.int32)
    mask = (((mb[:, None] >> sh[None, :]) & 1) != 0).reshape(-1)[:cap]
    ppe = torch.bincount(se[mask], minlength=m.nelems).to(torch.float64)
    ev = torch.as_tensor(m.elem2verts).cuda().long()
    acc = torch.zeros(m.nverts, dtype=torch.float64, device="cuda")
    for k in range(3):
        acc.index_add_(0, ev[:, k], ppe)
    pairs = (fmap.reshape(m.nverts, rings, ppr, 3)[:, :2] >= 0).sum(dim=(1, 2, 3)).to(torch.float64)
    expect = float((acc * pairs).sum().item()) / ppr
    got = float(pp.gyro_scatter(gm, ps, fmap, rmax, rings, ppr).sum().item())
    return {"particles": int(mask.sum().item()), "total_scattered": got, "total_expected": expect,
            "relative_error": abs(got - expect) / max(1.0, abs(expect))}


def c4x(pp, wl, torch, steps, nptcls, mdl_face, kind, fold_update=True):
    """BASELINE configs[3] on its named mesh: pseudoXGCm's step (test/pseudoXGCm.cpp:504-534) on
    pumipic-data/xgc/2M.osh (tests/golden/_large/mesh_xgc2M.npz, made by make_large_fixtures.py): the
    reference's class-limited normal particle load (:167-222) and initial coordinates (:224-264),
    ellipticalPush 0.5 degrees per step (h, k, d of :470-473), search_mesh_2d(maxLoops = 200),
    updatePtclPositions, rebuild, gyroScatter forward + backward (3 rings x 8 points, rmax 0.038)."""
    P = pp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    z = np.load(os.path.join(ROOT, "tests", "golden", "_large", "mesh_xgc2M.npz"))
    m = HostMesh()
    m.dim, m.coords, m.elem2verts = 2, z["coords"], z["elem2verts"]
    m.elem2sides, m.side2verts, m.class_id = z["elem2sides"], z["side2verts"], z["class_id_2"].astype(np.int32)
    m.nelems, m.nverts = m.elem2verts.shape[0], m.coords.shape[0]
    gm = pp.Mesh(2, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
    members = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float32, 1), (np.float32, 1)]   # pseudoXGCm_types.hpp
    ppe, total = wl.xgc_source_elements(m.class_id, np.zeros(m.nelems, np.int32), 0, mdl_face, nptcls)
    ps = pp.ParticleStructure(kind, members, ppe, V=1024, config={"extra_padding": 0.0})   # pseudoXGCm.cpp:452-465
    cap = ps.capacity
    slot_elem, mask = ps.slot_elem_and_mask()
    X = wl.xgc_initial_coords(m, slot_elem, mask)
    ps.get(0)[:, :cap] = torch.as_tensor(X).cuda()
    ps.get(2)[0, :cap] = torch.arange(cap, dtype=torch.int32, device="cuda")
    del X, slot_elem, mask
    h, k, d = 1.72479370 - .08, .020558260, 0.6
    P.elliptical_setup(ps, ps.get(0), ps.get(3), ps.get(4), h, k, d)
    rings, ppr, rmax = 3, 8, 0.038
    T = Timer(torch)
    fmap, st = T.run("gyro ring map (setup)", lambda: P.gyro_ring_map(gm, rmax, rings, ppr, 0.0))
    counts = [ps.nptcls]
    for it in range(steps + 1):
        ids = torch.full((ps.capacity,), -1, dtype=torch.int32, device="cuda")
        T.run("ellipticalPush", lambda: P.elliptical_push(gm, ps, ps.get(1), ps.get(3), ps.get(4), h, k, d, 0.5))
        T.run("search_mesh_2d", lambda: P.search_mesh(gm, ps, ps.get(0), ps.get(1), ids,
                                                       variant=P.capi.PP_SEARCH_2D_LEGACY, elem_ids_empty=True,
                                                       looplimit=200, sync=False))
        if fold_update:      # the reference's rebuild() starts with updatePtclPositions (pseudoXGCm.cpp:116-118):
            T.run("updatePtclPositions", lambda: ps.set_rebuild_remap([1, -1, 2, 3, 4]))   # folded into the record move
        else:
            T.run("updatePtclPositions", lambda: P.update_positions(ps, ps.get(0), ps.get(1)))
        T.run("rebuild", lambda: ps.rebuild(ids))
        T.run("gyroScatter x2", lambda: (P.gyro_scatter(gm, ps, fmap, rmax, rings, ppr),
                                         P.gyro_scatter(gm, ps, fmap, rmax, rings, ppr)))
        counts.append(ps.nptcls)
    s = T.summary()
    tot = sum(v["median_ms"] for k_, v in s.items() if "setup" not in k_)
    return {"config": "c4 on xgc/2M.osh", "particles": int(total), "triangles": m.nelems, "verts": m.nverts,
            "mdl_face": mdl_face, "elements_loaded": int((ppe > 0).sum()), "max_ppe": int(ppe.max()),
            "phases": s, "full_step_ms": tot, "full_steps_per_s": total / (tot * 1e-3),
            "particles_after": counts[-1], "capacity": ps.capacity,
            "updatePtclPositions": "member remap of the rebuild's record move" if fold_update else "own pass",
            "scatter_conservation_full_size": scatter_conservation(pp, torch, gm, ps, m, fmap, rmax, rings, ppr)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c3,c4")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--ps", default="scs")
    ap.add_argument("--c2-particles", type=int, default=10_000_000)
    ap.add_argument("--c4-particles", type=int, default=50_000_000)
    ap.add_argument("--mdl-face", type=int, default=426, help="c4x: particles go to elements with class id <= this")
    ap.add_argument("--rebuild-mode", type=int, default=2, help="pp_ps_set_staged_rebuild: 2 gather, 1 stage, 0 scatter")
    ap.add_argument("--chunk-order", type=int, default=1, help="pp_ps_set_rebuild_chunk_order")
    ap.add_argument("--shuffling", type=int, default=1, help="pp_ps_set_shuffling")
    ap.add_argument("--tuning", default="0,-1", help="pp_ps_set_rebuild_tuning: gather blocks/SM, gather max columns")
    ap.add_argument("--separate-update", action="store_true", help="c4x: updatePtclPositions as its own pass (A/B)")
    ap.add_argument("--block-histogram", type=int, default=1, help="pp_ps_set_rebuild_block_histogram (A/B)")
    a = ap.parse_args()
    import torch
    pp = importlib.import_module("pumi-pic_b200")
    wl = load_module("pp_workloads", os.path.join(ROOT, "pumi-pic_b200", "workloads.py"))
    kind = {"scs": pp.capi.PP_PS_SCS, "csr": pp.capi.PP_PS_CSR, "dps": pp.capi.PP_PS_DPS}[a.ps]
    pp.lib().pp_ps_set_staged_rebuild(a.rebuild_mode)
    pp.lib().pp_ps_set_rebuild_chunk_order(a.chunk_order)
    pp.lib().pp_ps_set_shuffling(a.shuffling)
    pp.lib().pp_ps_set_rebuild_block_histogram(a.block_histogram)
    pp.lib().pp_ps_set_rebuild_tuning(*[int(x) for x in a.tuning.split(",")])
    for c in a.configs.split(","):
        if c == "c2":
            r = c2(pp, wl, torch, a.steps, a.c2_particles, 55, kind)
        elif c == "c3":
            r = c3(pp, wl, torch, a.steps, 5500, 55_000_000 // 2, kind)
        elif c == "c4":
            r = c4(pp, wl, torch, a.steps, a.c4_particles, 1000, kind)
        elif c == "c4x":
            r = c4x(pp, wl, torch, a.steps, a.c4_particles, a.mdl_face, kind, fold_update=not a.separate_update)
        else:
            continue
        r["particle_structure"] = a.ps
        r["rebuild_mode"], r["chunk_order"], r["shuffling"] = a.rebuild_mode, a.chunk_order, a.shuffling
        r["tuning"] = a.tuning
        r["block_histogram"] = a.block_histogram
        print(json.dumps(r))
        sys.stdout.flush()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
