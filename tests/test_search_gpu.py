"""GPU parity: the fused CUDA search (through the C ABI) against the CPU oracle.

Element ids, intersection faces and counters must be bit-exact; intersection points and pushed
positions are compared bit-exactly too because the library is built with -fmad=false and
follows the reference's operation order (tolerance 0 stated here on purpose).
"""
import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from gpu_common import dev, make_gpu_mesh, make_ps, pp, torch
from meshes import kuhn_cube, load_fixture, plate

pytestmark = pytest.mark.gpu


def _setup(mesh, nptcls, kind=None):
    P = pp()
    kind = P.capi.PP_PS_DPS if kind is None else kind
    gm = make_gpu_mesh(mesh)
    ps = make_ps(kind, pi.even_ppe(mesh.nelems, nptcls))
    slot_elem, mask = ps.slot_elem_and_mask()
    init = pi.init3d_internal if mesh.dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    return gm, ps, slot_elem, mask, X, D


def test_mesh_derived_arrays_match_oracle():
    for mesh in (load_fixture("cube7k"), load_fixture("xgc24k"), kuhn_cube(5), plate(7)):
        om = orc.OracleMesh(mesh)
        gm = make_gpu_mesh(mesh)
        a = gm.arrays()
        assert np.array_equal(a["measure"], om.vol())            # bit-exact volumes
        assert np.array_equal(a["exposed"], om.exposed())
        off = om.side2elem_off(); s2e = om.side2elem()
        lo = s2e[off[:-1]]
        hi = np.where(np.diff(off) == 2, s2e[np.minimum(off[:-1] + 1, len(s2e) - 1)], -1)
        assert np.array_equal(a["side2elem"][:, 0], lo)
        assert np.array_equal(a["side2elem"][:, 1], hi)
        assert np.array_equal(a["dual_off"], om.dual_off())
        assert np.array_equal(a["dual"], om.dual())
        assert gm.info().tol == om.tol


@pytest.mark.parametrize("meshname,nptcls", [("cube7k", 20000), ("kuhn8", 30000),
                                             ("tri8", 100), ("xgc24k", 50000), ("plate20", 5000)])
def test_bcc_search_two_pushes(meshname, nptcls):
    """test_adj.cpp:746-781 testBCCSearch: push, search, push again, search again."""
    mesh = {"kuhn8": lambda: kuhn_cube(8), "plate20": lambda: plate(20)}.get(
        meshname, lambda: load_fixture(meshname))()
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, nptcls)
    dist = pi.push_distance(mesh)
    x = dev(X); tgt = dev(X.copy()); d = dev(D)
    ids = torch().full((ps.capacity,), -7, dtype=torch().int32, device="cuda")
    Xo, To = X.copy(), X.copy()
    ids_o = None
    for it in range(2):
        pp().push_direction(ps, tgt, d, dist)
        orc.push_direction(mask, To, D, dist)
        assert np.array_equal(tgt.cpu().numpy(), To)
        r = pp().search_mesh(gm, ps, x, tgt, ids, elem_ids_empty=(it == 0))
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, Xo, To, elem_ids=ids_o)
        got = ids.cpu().numpy()
        assert np.array_equal(got, ids_o)
        assert r.found == int(found) and r.loops == st.loops
        assert r.not_in_elem == st.not_in_elem == 0 and r.not_found == 0
        # test_parent_elements resets cur = tgt (test_adj.cpp:580-585)
        x.copy_(tgt); Xo[:] = To
    assert (ids_o[mask.astype(bool)] >= 0).mean() > 0.5


@pytest.mark.parametrize("meshname,nptcls", [("cube7k", 20000), ("kuhn8", 20000),
                                             ("tri8", 100), ("plate20", 5000)])
def test_intersection_search(meshname, nptcls):
    """test_adj.cpp:828-886: 10 pushes then ray search; walls give face + hit point."""
    mesh = {"kuhn8": lambda: kuhn_cube(8), "plate20": lambda: plate(20)}.get(
        meshname, lambda: load_fixture(meshname))()
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, nptcls)
    dist = pi.push_distance(mesh)
    T = X.copy()
    for _ in range(10):
        orc.push_direction(mask, T, D, dist)
    x = dev(X); tgt = dev(T)
    cap = ps.capacity
    t = torch()
    ids = t.zeros(cap, dtype=t.int32, device="cuda")
    faces = t.full((cap,), 5, dtype=t.int32, device="cuda")
    pts = t.full((cap * mesh.dim,), 3.0, dtype=t.float64, device="cuda")
    r = pp().search_mesh(gm, ps, x, tgt, ids, elem_ids_empty=True, require_intersection=True,
                         inter_faces=faces, inter_points=pts)
    found, ids_o, faces_o, pts_o, st = om.search_mesh(slot_elem, mask, X, T,
                                                      require_intersection=True)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert np.array_equal(faces.cpu().numpy(), faces_o)
    assert np.array_equal(pts.cpu().numpy().reshape(cap, mesh.dim), pts_o)   # bit-exact
    assert r.found == int(found) and r.loops == st.loops
    assert (faces_o >= 0).sum() > 0


def test_search_deletes_particles_outside_parent_and_honours_looplimit():
    mesh = load_fixture("cube7k")
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 5000)
    # move every 11th particle to the centroid of a far element => not in its parent
    bad = np.arange(0, mask.shape[0], 11)
    far = (slot_elem[bad] + mesh.nelems // 2) % mesh.nelems
    X[:, bad] = mesh.coords[mesh.elem2verts[far]].mean(axis=1).T
    T = X.copy()
    orc.push_direction(mask, T, D, 25.0)     # long push: many hops
    t = torch()
    for limit in (0, 3):
        ids = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
        r = pp().search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True, looplimit=limit)
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, X, T, looplimit=limit)
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert (r.found, r.loops, r.not_in_elem, r.not_found) == \
               (int(found), st.loops, st.not_in_elem, st.not_found)
        assert r.not_in_elem > 0
    assert r.not_found > 0 and not r.found


def test_search_mesh_2d_legacy_matches_oracle_and_goldens():
    from test_oracle_golden import TRI8_CASES
    mesh = load_fixture("tri8_parDiag")
    gm = make_gpu_mesh(mesh)
    t = torch()
    P = pp()
    for parent, start, end, dest, alt in TRI8_CASES:       # test/search2d.cpp:186-309
        ppe = np.zeros(mesh.nelems, np.int32); ppe[parent] = 1
        ps = make_ps(P.capi.PP_PS_DPS, ppe)
        X = np.zeros((3, ps.capacity)); T = np.zeros((3, ps.capacity))
        X[:2, 0] = start; T[:2, 0] = end
        ids = t.full((ps.capacity,), -1, dtype=t.int32, device="cuda")
        r = P.search_mesh(gm, ps, dev(X), dev(T), ids, variant=P.capi.PP_SEARCH_2D_LEGACY,
                          looplimit=100)
        assert r.found
        assert int(ids[0]) in (dest, alt)
    mesh = load_fixture("xgc24k")
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 40000)
    T = X.copy()
    orc.push_direction(mask, T, D, 4 * pi.push_distance(mesh))
    start_ids = np.full(ps.capacity, -1, np.int32)
    start_ids[::13] = -mesh.nelems                     # "already outside" sentinel (hpp:1053-1056)
    ids = dev(start_ids)
    r = P.search_mesh(gm, ps, dev(X), dev(T), ids, variant=P.capi.PP_SEARCH_2D_LEGACY, looplimit=200)
    found, ids_o, st = om.search_mesh_2d(slot_elem, mask, T, start_ids, looplimit=200)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert (r.found, r.loops) == (int(found), st.loops)


def test_legacy_3d_search_matches_oracle():
    """adjacency.hpp:559 as driven by test/pseudoPushAndSearch.cpp: constant push in +z."""
    mesh = load_fixture("cube7k")
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 20000)
    ext = (mesh.coords.max(axis=0) - mesh.coords.min(axis=0)).max()
    T = np.zeros_like(X)
    orc.push_constant(mask, X, T, ext / 20, (0.0, 0.0, 1.0))   # pseudoPushAndSearch.cpp:482-496
    t = torch()
    P = pp()
    x = dev(X); tg = t.zeros_like(x)
    P.push_constant(ps, x, tg, ext / 20, (0.0, 0.0, 1.0))
    assert np.array_equal(tg.cpu().numpy(), T)
    cap = ps.capacity
    ids = t.zeros(cap, dtype=t.int32, device="cuda")
    xface = t.full((cap,), -1, dtype=t.int32, device="cuda")
    xpts = t.zeros(3 * cap, dtype=t.float64, device="cuda")
    r = P.search_mesh(gm, ps, x, tg, ids, elem_ids_empty=True, variant=P.capi.PP_SEARCH_3D_LEGACY,
                      inter_faces=xface, inter_points=xpts, looplimit=100)
    found, ids_o, xp_o, xf_o, st = om.search_mesh_legacy3d(slot_elem, mask, X, T, looplimit=100)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert np.array_equal(xface.cpu().numpy(), xf_o)
    assert np.array_equal(xpts.cpu().numpy().reshape(cap, 3), xp_o)
    assert (r.found, r.loops, r.aborted) == (int(found), st.loops, st.aborted)
    assert (xf_o >= 0).sum() > 0 and (ids_o >= 0).sum() > 0


@pytest.mark.parametrize("meshname,push", [("cube7k", "constant"), ("cube7k", "direction"),
                                           ("kuhn8", "direction")])
def test_search_mesh_3d_matches_oracle(meshname, push):
    """adjacency.hpp:316 search_mesh_3d (PP_SEARCH_3D): element ids, wall faces, wall points and
    counters bit-exact against the oracle, two consecutive steps, with and without a loop limit."""
    mesh = kuhn_cube(8) if meshname == "kuhn8" else load_fixture(meshname)
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 30000)
    ext = (mesh.coords.max(axis=0) - mesh.coords.min(axis=0)).max()
    t = torch()
    P = pp()
    cap = ps.capacity
    T = np.zeros_like(X)
    if push == "constant":
        orc.push_constant(mask, X, T, ext / 20, (0.0, 0.0, 1.0))
    else:
        T = X.copy()
        orc.push_direction(mask, T, D, 6 * pi.push_distance(mesh))
    for limit in (0, 3):
        ids = t.zeros(cap, dtype=t.int32, device="cuda")
        xface = t.full((cap,), -1, dtype=t.int32, device="cuda")
        xpts = t.zeros(3 * cap, dtype=t.float64, device="cuda")
        r = P.search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True, variant=P.capi.PP_SEARCH_3D,
                          inter_faces=xface, inter_points=xpts, looplimit=limit)
        found, ids_o, xp_o, xf_o, st = om.search_mesh_3d(slot_elem, mask, X, T, looplimit=limit)
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert np.array_equal(xface.cpu().numpy(), xf_o)
        assert np.array_equal(xpts.cpu().numpy().reshape(cap, 3), xp_o)
        assert (r.found, r.loops, r.aborted, r.not_found) == (int(found), st.loops, st.aborted, st.not_found)
        if limit == 0:
            assert found and (xf_o >= 0).sum() > 0 and (ids_o >= 0).sum() > 0
        else:
            assert not found and st.not_found > 0
    # second step from the found elements (elem_ids passed in), origin = previous target
    keep = ids_o.copy()
    T2 = T.copy()
    orc.push_direction(mask, T2, D, 2 * pi.push_distance(mesh))
    ids = dev(keep)
    xface = t.full((cap,), -1, dtype=t.int32, device="cuda")
    xpts = t.zeros(3 * cap, dtype=t.float64, device="cuda")
    r = P.search_mesh(gm, ps, dev(T), dev(T2), ids, elem_ids_empty=False, variant=P.capi.PP_SEARCH_3D,
                      inter_faces=xface, inter_points=xpts, looplimit=100)
    found, ids_o2, xp_o2, xf_o2, st2 = om.search_mesh_3d(slot_elem, mask, T, T2, elem_ids=keep, looplimit=100)
    assert np.array_equal(ids.cpu().numpy(), ids_o2)
    assert np.array_equal(xface.cpu().numpy(), xf_o2)
    assert np.array_equal(xpts.cpu().numpy().reshape(cap, 3), xp_o2)
    assert (r.found, r.loops, r.aborted) == (int(found), st2.loops, st2.aborted)


def test_fused_push_search_equals_push_then_search():
    mesh = kuhn_cube(8)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 25000)
    dist = pi.push_distance(mesh)
    t = torch()
    P = pp()
    x = dev(X); d = dev(D)
    tg1 = dev(X.copy()); tg2 = dev(X.copy())
    ids1 = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
    ids2 = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
    for it in range(3):
        P.push_direction(ps, tg1, d, dist)
        r1 = P.search_mesh(gm, ps, x, tg1, ids1, elem_ids_empty=(it == 0))
        r2 = P.push_direction_search(gm, ps, d, dist, x, tg2, ids2, elem_ids_empty=(it == 0))
        assert t.equal(tg1, tg2) and t.equal(ids1, ids2)
        assert (r1.found, r1.loops, r1.hops, r1.active) == (r2.found, r2.loops, r2.hops, r2.active)
        x.copy_(tg1)


def test_fused_pic_form_push_search_ping_pong():
    """xtgt = x + d*dir fused with the walk, buffers swapped every step (bench.py's step)."""
    mesh = kuhn_cube(8)
    om = orc.OracleMesh(mesh)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 25000)
    dist = pi.push_distance(mesh)
    t = torch()
    P = pp()
    a = dev(X); b = t.zeros_like(a); d = dev(D)
    ids = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
    A, B = X.copy(), np.zeros_like(X)
    ids_o = None
    for it in range(4):
        sgn = dist if it % 2 == 0 else -dist
        r = P.push_direction_search(gm, ps, d, sgn, a, b, ids, elem_ids_empty=(it == 0),
                                    from_orig=True)
        m = mask.astype(bool)
        B[:, m] = A[:, m] + sgn * D[:, m]
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, A, B, elem_ids=ids_o)
        assert np.array_equal(b.cpu().numpy()[:, m], B[:, m])
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert (r.found, r.loops, r.not_in_elem) == (int(found), st.loops, st.not_in_elem)
        # unfused twin
        b2 = t.zeros_like(a)
        P.push_from(ps, a, b2, d, sgn)
        assert t.equal(b2[:, t.as_tensor(m).cuda()], b[:, t.as_tensor(m).cuda()])
        a, b = b, a
        A, B = B, A


def test_update_positions_ignores_mask():
    mesh = kuhn_cube(3)
    gm, ps, slot_elem, mask, X, D = _setup(mesh, 500)
    t = torch()
    x = t.rand(3, ps.capacity, dtype=t.float64, device="cuda")
    tg = t.rand(3, ps.capacity, dtype=t.float64, device="cuda")
    want = tg.clone()
    pp().update_positions(ps, x, tg)
    assert t.equal(x, want) and float(tg.abs().max()) == 0.0


def test_boris_push_matches_oracle_bitwise():
    """src/pumipic_push.hpp:17-74: three successive Boris steps, positions/velocities bit-exact."""
    rng = np.random.default_rng(5)
    n = 10007
    pos = rng.normal(size=(3, n)); prev = pos - 1e-3 * rng.normal(size=(3, n))
    vel = 1e4 * rng.normal(size=(3, n)); E = 50.0 * rng.normal(size=(3, n)); B = rng.normal(size=(3, n))
    B[:, ::97] = 0.0
    t = torch()
    dpos, dprev, dvel, dE, dB = (dev(a.copy()) for a in (pos, prev, vel, E, B))
    for _ in range(3):
        orc.push_boris(pos, prev, vel, E, B, 1e-9)
        pp().push_boris(dpos, dprev, dvel, dE, dB, 1e-9)
    assert np.array_equal(dpos.cpu().numpy(), pos)
    assert np.array_equal(dprev.cpu().numpy(), prev)
    assert np.array_equal(dvel.cpu().numpy(), vel)
