#!/usr/bin/env python
"""Diagnostic: search_mesh_2d on the pseudoXGCm load (xgc/2M.osh) before and after a rebuild, with the walk's
counters and the structure's slice / chunk counts; chunk walk vs block-staged kernel."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_phases as bp
pp = importlib.import_module("pumi-pic_b200")
wl = bp.load_module("pp_workloads", os.path.join(ROOT, "pumi-pic_b200", "workloads.py"))
P = pp
z = np.load(os.path.join(ROOT, "tests", "golden", "_large", "mesh_xgc2M.npz"))
m = bp.HostMesh()
m.dim, m.coords, m.elem2verts = 2, z["coords"], z["elem2verts"]
m.elem2sides, m.side2verts, m.class_id = z["elem2sides"], z["side2verts"], z["class_id_2"].astype(np.int32)
m.nelems, m.nverts = m.elem2verts.shape[0], m.coords.shape[0]
gm = pp.Mesh(2, m.coords, m.elem2verts, m.elem2sides, m.side2verts, m.class_id)
members = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float32, 1), (np.float32, 1)]
nptcls = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
ppe, total = wl.xgc_source_elements(m.class_id, np.zeros(m.nelems, np.int32), 0, 426, nptcls)
ps = pp.ParticleStructure(pp.capi.PP_PS_SCS, members, ppe, V=1024, config={"extra_padding": 0.0})
cap = ps.capacity
slot_elem, mask = ps.slot_elem_and_mask()
X = wl.xgc_initial_coords(m, slot_elem, mask)
ps.get(0)[:, :cap] = torch.as_tensor(X).cuda()
h, k, d = 1.72479370 - .08, .020558260, 0.6
P.elliptical_setup(ps, ps.get(0), ps.get(3), ps.get(4), h, k, d)

def search(tag, staged):
    P.lib().pp_search_set_staged(staged)
    lay = ps.layout()
    ids = torch.full((ps.capacity,), -1, dtype=torch.int32, device="cuda")
    P.elliptical_push(gm, ps, ps.get(1), ps.get(3), ps.get(4), h, k, d, 0.5)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    st = P.search_mesh(gm, ps, ps.get(0), ps.get(1), ids, variant=P.capi.PP_SEARCH_2D_LEGACY, elem_ids_empty=True,
                       looplimit=200, sync=True)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"tag": tag, "staged": staged, "ms": a.elapsed_time(b), "nchunks": lay.nchunks, "nslices": lay.nslices,
                      "capacity": lay.capacity, "nptcls": lay.nptcls, "stats": repr(st)}))
    sys.stdout.flush()
    return ids

ids = search("built", 2)
# undo the phase advance of the diagnostic push so that the next push is one step again: not needed for timing
P.update_positions(ps, ps.get(0), ps.get(1))
ps.rebuild(ids)
ids = search("after rebuild (fast path)", 2)
ids = search("after rebuild, block-staged kernel", 1)
P.update_positions(ps, ps.get(0), ps.get(1))
P.lib().pp_ps_set_staged_rebuild(1)
ps.rebuild(ids)
ids = search("after rebuild (mode 1)", 2)
