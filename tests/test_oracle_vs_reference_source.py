"""CPU-only: the oracle against THE REFERENCE'S OWN CODE, bit for bit.

oracle/build_ref_primitives.py compiles, unmodified and straight from /root/reference,
  * the geometric primitives: barycentric_tet / barycentric_tri / ray_intersects_triangle /
    line_segment_intersects_triangle / line_edge_2d / find_exit_face_bcc_3d
    (src/pumipic_adjacency.tpp:23-228), find_barycentric_tet / find_barycentric_tri_simple /
    line_triangle_intx_simple (src/pumipic_adjacency.hpp:97-273), all_positive / min3 / min_index /
    max_index / isFaceFlipped (src/pumipic_utils.hpp:78-149,489-507);
  * the search loops: search_mesh -> trace_particle_through_mesh and its kernels
    (adjacency.tpp:72-660), search_mesh_2d (adjacency.hpp:1013-1158), the legacy 3D search_mesh
    (:559-768) and search_mesh_3d (:316-555)
against stand-ins for the Omega_h / Kokkos vocabulary (oracle/ref_shim/).  Every restatement in
oracle/pumipic_oracle.c must return exactly the same doubles, ids and decisions.  What this does
NOT pin is Omega_h's own arithmetic and mesh derivations (cross, inner_product, ask_up ...): the
shims restate them from the published definitions, as the oracle does.  Skipped where the library
was never built (no /root/reference and no prebuilt copy).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_api as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpumipic_ref_primitives.so")
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def ref():
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref_primitives.py")],
                          stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB):
        pytest.skip("reference primitives not built (needs /root/reference once)")
    return C.CDLL(LIB)


def _d(a):
    return np.ascontiguousarray(a, np.float64).ctypes.data_as(dp)


def _i(a):
    return np.ascontiguousarray(a, np.int32).ctypes.data_as(ip)


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def _tets(rng, n):
    """random tets, some flat or inverted, and points inside / on faces / on vertices / far away"""
    M = rng.random((n, 4, 3))
    M[::7, 3] = M[::7, 0] + 1e-14 * rng.random((len(M[::7]), 3))       # nearly degenerate
    M[::11, 3] = M[::11, 0]                                              # degenerate
    w = rng.random((n, 4)); w /= w.sum(axis=1, keepdims=True)
    p = np.einsum("nk,nkd->nd", w, M)
    p[::3] += rng.normal(0, 0.5, p[::3].shape)                           # outside
    p[1::5] = M[1::5, 1]                                                 # on a vertex
    p[2::5] = 0.5 * (M[2::5, 0] + M[2::5, 2])                            # on an edge
    return M, p


def test_barycentric_and_exit_face(ref):
    L = orc.lib()
    rng = np.random.default_rng(1)
    M, P = _tets(rng, 4000)
    for m, p in zip(M, P):
        e1, e2, e3 = m[1] - m[0], m[2] - m[0], m[3] - m[0]
        vol = float(np.dot(np.cross(e1, e2), e3) / 6.0)
        for v in (vol, -vol, 0.0):
            a, b = np.zeros(4), np.zeros(4)
            ra = L.orc_barycentric_tet(C.c_double(v), _d(m), _d(p), _d(a))
            rb = ref.ref_barycentric_tet(C.c_double(v), _d(m), _d(p), _d(b))
            assert ra == rb and _same(a, b)
        a, b = np.zeros(4), np.zeros(4)
        assert L.orc_find_barycentric_tet(_d(m), _d(p), _d(a)) == ref.ref_find_barycentric_tet(_d(m), _d(p), _d(b))
        assert _same(a, b)
        # find_exit_face_bcc_3d = barycentric_tet + all_positive(EPSILON) + min_index
        done = C.c_int()
        f = ref.ref_find_exit_face_bcc_3d(C.c_double(vol), _d(m), _d(p), C.byref(done))
        L.orc_barycentric_tet(C.c_double(vol), _d(m), _d(p), _d(a))
        assert f == L.orc_min_index(_d(a), 4) and done.value == L.orc_all_positive(_d(a), 4, C.c_double(1e-10))
    # triangles
    T = rng.random((3000, 3, 2))
    Q = rng.random((3000, 2)) * 1.5 - 0.25
    Q[::4] = T[::4, 1]
    for t, q in zip(T, Q):
        area = 0.5 * float((t[1, 0] - t[0, 0]) * (t[2, 1] - t[0, 1]) - (t[1, 1] - t[0, 1]) * (t[2, 0] - t[0, 0]))
        a, b = np.zeros(3), np.zeros(3)
        L.orc_barycentric_tri(C.c_double(area), _d(t), _d(q), _d(a))
        ref.ref_barycentric_tri(C.c_double(area), _d(t), _d(q), _d(b))
        assert _same(a, b)
        assert L.orc_min3(_d(a)) == ref.ref_min3(_d(b))


def test_all_positive_min_max_index(ref):
    L = orc.lib()
    rng = np.random.default_rng(2)
    special = [0.0, -0.0, 1e-10, -1e-10, np.nextafter(-1e-10, -1), np.nextafter(-1e-10, 0), 1e-8, -1e-8,
               np.inf, -np.inf, np.nan, 1.0, -1.0, 5e-11, -5e-11]
    for k in range(6000):
        n = 3 if k % 2 else 4
        v = rng.normal(0, 1e-9 if k % 3 else 1.0, n)
        if k % 5 == 0:
            v[rng.integers(0, n)] = special[k // 5 % len(special)]
        if k % 7 == 0:
            v[1] = v[0]                                               # ties
        for tol in (1e-10, 1e-8, 1e-20, 0.0, 1.0, 2.5):
            assert L.orc_all_positive(_d(v), n, C.c_double(tol)) == ref.ref_all_positive(_d(v), n, C.c_double(tol)), (v, tol)
        if not np.isnan(v).any():
            assert L.orc_min_index(_d(v), n) == ref.ref_min_index(_d(v), n)
            assert L.orc_max_index(_d(v), n) == ref.ref_max_index(_d(v), n)
            if n == 3:
                assert L.orc_min3(_d(v)) == ref.ref_min3(_d(v))


def test_face_flips(ref):
    L = orc.lib()
    import itertools
    tv = np.array([11, 5, 7, 3], np.int32)
    faces = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)]
    for fi, f in enumerate(faces):
        for perm in itertools.permutations(f):
            fv = tv[list(perm)]
            assert L.orc_is_face_flipped_3d(fi, _i(fv), _i(tv)) == ref.ref_is_face_flipped_3d(fi, _i(fv), _i(tv))
    t3 = np.array([4, 9, 2], np.int32)
    for ei, e in enumerate([(0, 1), (1, 2), (2, 0)]):
        for perm in itertools.permutations(e):
            ev = t3[list(perm)]
            assert L.orc_is_face_flipped_2d(_i(ev), _i(t3)) == ref.ref_is_face_flipped_2d(ei, _i(ev), _i(t3))


def test_ray_segment_edge_and_legacy_intersections(ref):
    L = orc.lib()
    rng = np.random.default_rng(3)
    for k in range(5000):
        face = rng.random((3, 3))
        o = rng.random(3) * 2 - 0.5
        d = rng.random(3) * 2 - 0.5
        if k % 6 == 0:                                                # ray in the face's plane
            d = o + (face[1] - face[0])
        if k % 9 == 0:                                                # through a vertex
            d = o + 2 * (face[2] - o)
        if k % 50 == 0:
            d = o.copy()                                              # zero-length path: 0/0
        for flip in (0, 1):
            for tol in (1e-8, 0.0):
                out = []
                for fn_ray, fn_seg in ((L.orc_ray_intersects_triangle, L.orc_line_segment_intersects_triangle),
                                       (ref.ref_ray_intersects_triangle, ref.ref_line_segment_intersects_triangle)):
                    xp, xs = np.zeros(3), np.zeros(3)
                    a, b, c = C.c_double(), C.c_double(), C.c_double()
                    a2, b2, c2 = C.c_double(), C.c_double(), C.c_double()
                    r1 = fn_ray(_d(face), _d(o), _d(d), _d(xp), C.c_double(tol), flip, C.byref(a), C.byref(b), C.byref(c))
                    r2 = fn_seg(_d(face), _d(o), _d(d), _d(xs), C.c_double(tol), flip, C.byref(a2), C.byref(b2), C.byref(c2))
                    out.append((r1, r2, xp, xs, [a.value, b.value, c.value, a2.value, b2.value, c2.value]))
                assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
                assert _same(out[0][2], out[1][2]) and _same(out[0][3], out[1][3]) and _same(out[0][4], out[1][4])
        for reverse in (0, 1):
            res = []
            for fn in (L.orc_line_triangle_intx_simple, ref.ref_line_triangle_intx_simple):
                xp = np.zeros(3)
                dpj = C.c_double(-7.0)
                r = fn(_d(face), _d(o), _d(d), _d(xp), C.byref(dpj), reverse, C.c_double(1e-10))
                res.append((r, xp, dpj.value))
            assert res[0][0] == res[1][0] and _same(res[0][1], res[1][1]) and _same(res[0][2], res[1][2])
        # 2D segment against an edge
        e = rng.random(4)
        o2, d2 = rng.random(2) * 2 - 0.5, rng.random(2) * 2 - 0.5
        if k % 8 == 0:
            d2 = o2 + (e[2:] - e[:2])                                 # parallel to the edge
        for flip in (0, 1):
            xa, xb = np.zeros(2), np.zeros(2)
            ra = L.orc_line_edge_2d(_d(e), _d(o2), _d(d2), _d(xa), C.c_double(1e-8), flip)
            rb = ref.ref_line_edge_2d(_d(e), _d(o2), _d(d2), _d(xb), C.c_double(1e-8), flip)
            assert ra == rb and _same(xa, xb)


# ---------------------------------------------------------------- the search loops themselves
def _ref_search(ref, om, mesh, slot_elem, mask, X, T, elem_ids=None, require_intersection=False,
                looplimit=0, inter_given=False):
    cap = mask.shape[0]
    dim = mesh.dim
    ids = np.full(cap, -1, np.int32) if elem_ids is None else np.ascontiguousarray(elem_ids, np.int32).copy()
    faces = np.full(cap, 7, np.int32) if inter_given else np.full(cap, -1, np.int32)
    pts = np.full(dim * cap, 2.5) if inter_given else np.zeros(dim * cap)
    off, val = om.side2elem_off(), om.side2elem()
    X, T = np.ascontiguousarray(X, np.float64), np.ascontiguousarray(T, np.float64)
    found = ref.ref_search_mesh(
        dim, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides,
        _i(mesh.elem2sides), _i(mesh.side2verts), _i(off), _i(val),
        np.ascontiguousarray(om.exposed(), np.int8).ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()), cap,
        _i(slot_elem), np.ascontiguousarray(mask, np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte)),
        _d(X), _d(T), C.c_long(X.shape[1]), ids.ctypes.data_as(ip), int(elem_ids is None),
        int(require_intersection), faces.ctypes.data_as(ip), pts.ctypes.data_as(dp), int(inter_given), looplimit)
    return bool(found), ids, faces, pts.reshape(cap, dim)


@pytest.mark.parametrize("meshname,n,mult", [("kuhn5", 5000, 5.0), ("cube7k", 12000, 3.0), ("plate15", 3000, 5.0),
                                             ("xgc24k", 8000, 6.0), ("tri8", 300, 2.0)])
def test_search_mesh_equals_the_reference_loops(ref, meshname, n, mult):
    """The oracle's search_mesh against the reference's own search_mesh -> trace_particle_through_mesh
    (adjacency.tpp:461-660, with its find_exit_face, check_model_intersection, set_new_element,
    check_initial_parents and compute_tolerance_from_area) compiled unmodified and run serially over
    the same mesh arrays: element ids, wall sides and wall points must be identical, in both modes,
    with fresh and carried-over element ids, deleted particles, wrong parents and a loop limit."""
    import ptcl_init as pi
    from meshes import kuhn_cube, load_fixture, plate
    mesh = {"kuhn5": lambda: kuhn_cube(5), "plate15": lambda: plate(15)}.get(meshname, lambda: load_fixture(meshname))()
    om = orc.OracleMesh(mesh)
    slot_elem = ((np.arange(n, dtype=np.int64) * 7919) % mesh.nelems).astype(np.int32)
    mask = np.ones(n, np.uint8)
    mask[::13] = 0
    init = pi.init3d_internal if mesh.dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    m = mask.astype(bool)
    T = X.copy()
    T[:, m] = X[:, m] + mult * pi.push_distance(mesh) * D[:, m]
    T[:, 5::41] = X[:, 5::41]                                   # particles that do not move
    for req in (False, True):
        f0, i0, x0, p0, st = om.search_mesh(slot_elem, mask, X, T, require_intersection=req)
        f1, i1, x1, p1 = _ref_search(ref, om, mesh, slot_elem, mask, X, T, require_intersection=req)
        assert f0 == f1 and np.array_equal(i0, i1)
        if req:
            assert np.array_equal(x0, x1) and _same(p0, p1) and (x0 >= 0).any()
        # carried-over ids with deletions and wrong parents, arrays handed in dirty, and a loop limit
        start = np.where(m, slot_elem, -1).astype(np.int32)
        live = np.flatnonzero(m)
        start[live[::11]] = -1
        start[live[3::17]] = (slot_elem[live[3::17]] + mesh.nelems // 2) % mesh.nelems
        for limit in (0, 2):
            f0, i0, x0, p0, st = om.search_mesh(slot_elem, mask, X, T, elem_ids=start, require_intersection=req,
                                                looplimit=limit)
            f1, i1, x1, p1 = _ref_search(ref, om, mesh, slot_elem, mask, X, T, elem_ids=start,
                                         require_intersection=req, looplimit=limit, inter_given=req)
            assert f0 == f1 and np.array_equal(i0, i1) and st.not_in_elem > 0
            if req:
                assert np.array_equal(x0, x1) and _same(p0, p1)


@pytest.mark.parametrize("meshname,n,mult", [("plate15", 3000, 5.0), ("xgc24k", 8000, 6.0), ("tri8", 300, 2.0),
                                             ("tri8_parDiag", 300, 3.0)])
def test_search_mesh_2d_equals_the_reference_loop(ref, meshname, n, mult):
    """The oracle's search_mesh_2d against the reference's own (adjacency.hpp:1013-1158, with the
    5-argument barycentric_tri of :75-94) compiled unmodified: identical element ids, including the
    `-nelems` sentinel, carried-over ids and the loop limit."""
    import ptcl_init as pi
    from meshes import load_fixture, plate
    mesh = {"plate15": lambda: plate(15)}.get(meshname, lambda: load_fixture(meshname))()
    om = orc.OracleMesh(mesh)
    slot_elem = ((np.arange(n, dtype=np.int64) * 7919) % mesh.nelems).astype(np.int32)
    mask = np.ones(n, np.uint8)
    mask[::13] = 0
    X, D = pi.init2d_internal(mesh, slot_elem, mask)
    m = mask.astype(bool)
    T = X.copy()
    T[:, m] = X[:, m] + mult * pi.push_distance(mesh) * D[:, m]
    off, val = om.side2elem_off(), om.side2elem()
    start = np.full(n, -1, np.int32)
    live = np.flatnonzero(m)
    start[live[::7]] = slot_elem[live[::7]]                      # some given, some -1 (row element)
    start[live[2::19]] = -mesh.nelems                            # "already outside" sentinel (:1053-1056)
    for limit in (0, 2):
        f0, i0, st = om.search_mesh_2d(slot_elem, mask, T, start, looplimit=limit)
        ids = start.copy()
        f1 = ref.ref_search_mesh_2d(
            mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides, _i(mesh.elem2sides),
            _i(mesh.side2verts), _i(off), _i(val),
            np.ascontiguousarray(om.exposed(), np.int8).ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()), n,
            _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(X), _d(T), C.c_long(X.shape[1]),
            ids.ctypes.data_as(ip), limit)
        assert f0 == bool(f1) and np.array_equal(i0, ids)
        assert (ids >= 0).any() and (ids[m] == -1).any()


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("meshname,n,mult", [("kuhn5", 4000, 4.0), ("cube7k", 10000, 3.0)])
def test_legacy_3d_searches_equal_the_reference_loops(ref, meshname, n, mult, variant):
    """variant 0: the legacy 3D search_mesh (adjacency.hpp:559-768); variant 1: search_mesh_3d
    (:316-555, which no reference test exercises).  Both compiled unmodified -- including the
    dual-graph-indexed-by-face-id fallbacks (:726, :510) -- and compared with the oracle's
    restatements: element ids, wall points and wall faces, with fresh and carried-over ids and loop
    limits; where the reference would abort (origin outside the start element) the oracle must
    report it."""
    import ptcl_init as pi
    from meshes import kuhn_cube, load_fixture
    mesh = kuhn_cube(5) if meshname == "kuhn5" else load_fixture(meshname)
    om = orc.OracleMesh(mesh)
    slot_elem = ((np.arange(n, dtype=np.int64) * 7919) % mesh.nelems).astype(np.int32)
    mask = np.ones(n, np.uint8)
    mask[::13] = 0
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    m = mask.astype(bool)
    T = X.copy()
    T[:, m] = X[:, m] + mult * pi.push_distance(mesh) * D[:, m]
    off, val, doff, dval = om.side2elem_off(), om.side2elem(), om.dual_off(), om.dual()
    oracle_fn = om.search_mesh_legacy3d if variant == 0 else om.search_mesh_3d

    def run_ref(elem_ids, limit):
        ids = np.full(n, -1, np.int32) if elem_ids is None else elem_ids.copy()
        xp, xf = np.zeros(3 * n), np.full(n, -1, np.int32)
        r = ref.ref_search_mesh_3d_variants(
            variant, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides,
            _i(mesh.elem2sides), _i(mesh.side2verts), _i(off), _i(val),
            np.ascontiguousarray(om.exposed(), np.int8).ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()),
            _i(doff), _i(dval), n, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(X), _d(T),
            C.c_long(X.shape[1]), ids.ctypes.data_as(ip), int(elem_ids is None), xp.ctypes.data_as(dp),
            xf.ctypes.data_as(ip), limit)
        return r, ids, xp.reshape(n, 3), xf

    carried = np.where(m, slot_elem, -1).astype(np.int32)
    carried[np.flatnonzero(m)[::9]] = -1
    for start in (None, carried):
        for limit in (40, 3):
            f0, i0, p0, x0, st = oracle_fn(slot_elem, mask, X, T, elem_ids=start, looplimit=limit)
            r, i1, p1, x1 = run_ref(start, limit)
            assert st.aborted == 0 and r == int(f0)
            assert np.array_equal(i0, i1) and np.array_equal(x0, x1) and _same(p0, p1)
            assert (x0 >= 0).any() and (i0 >= 0).any()
    # a particle whose origin is not in its start element: the reference aborts, the oracle says so
    bad = carried.copy()
    k = np.flatnonzero(m & (carried >= 0))[5]
    bad[k] = (slot_elem[k] + mesh.nelems // 2) % mesh.nelems
    if variant == 0:                                   # legacy: checks the element it starts the walk in (:619-627)
        f0, i0, p0, x0, st = oracle_fn(slot_elem, mask, X, T, elem_ids=bad, looplimit=40)
        assert st.aborted >= 1 and run_ref(bad, 40)[0] == -2
    else:                                              # search_mesh_3d: checks the ROW element (:368-379)
        f0, i0, p0, x0, st = oracle_fn(slot_elem, mask, X, T, elem_ids=bad, looplimit=40)
        r, i1, p1, x1 = run_ref(bad, 40)
        assert st.aborted == 0 and r == int(f0) and np.array_equal(i0, i1)


@pytest.mark.parametrize("meshname,rmax,nrings,ppr,theta", [("tri8_parDiag", 0.2, 2, 6, 15), ("plate15", 0.11, 3, 8, 0),
                                                            ("xgc24k", 0.038, 3, 8, 0), ("xgc24k", 0.05, 4, 5, 7)])
def test_gyro_ring_map_and_scatter_equal_the_reference(ref, meshname, rmax, nrings, ppr, theta):
    """createGyroRingMappings + searchAndBuildMap (test/gyroScatter.hpp:25-166, which builds a throw-away
    Sell-C-sigma of ring points and runs search_mesh_2d on it) and gyroScatter (:168-229), compiled
    unmodified: the oracle's ring map must be identical, and its scatter equal to the last bit (the
    addends are small integers / points-per-ring, summed in the same vertex order)."""
    from meshes import load_fixture, plate
    mesh = plate(15) if meshname == "plate15" else load_fixture(meshname)
    om = orc.OracleMesh(mesh)
    off, val = om.side2elem_off(), om.side2elem()
    voff, vval = om.vert2elem_off(), om.vert2elem()
    npts = mesh.nverts * nrings * ppr
    fwd = np.full(3 * npts, -9, np.int32)
    same = C.c_int()
    n = ref.ref_gyro_ring_map(
        mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides, _i(mesh.elem2sides),
        _i(mesh.side2verts), _i(off), _i(val),
        np.ascontiguousarray(om.exposed(), np.int8).ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()),
        _i(voff), _i(vval), C.c_double(rmax), nrings, ppr, theta, fwd.ctypes.data_as(ip), C.byref(same))
    assert n == 3 * npts and same.value == 1
    found, mine = om.gyro_ring_map(rmax, nrings, ppr, float(theta))
    assert found and np.array_equal(mine, fwd)
    assert (fwd >= 0).any() and (meshname == "xgc24k" or (fwd == -1).any())      # ring points outside the domain
    # scatter: particles on a skewed subset of the elements, some slots empty
    rng = np.random.default_rng(4)
    cap = 5000
    slot_elem = rng.integers(0, mesh.nelems, cap).astype(np.int32)
    slot_elem[: cap // 3] = slot_elem[0]                          # a crowded element
    mask = (rng.random(cap) < 0.8).astype(np.uint8)
    got = om.gyro_scatter(slot_elem, mask, fwd, rmax, nrings, ppr)
    nthreads = ref.ref_get_max_threads()
    for threads in (1, nthreads):
        # the reference sums with atomics: one thread = ascending vertex order (the oracle's order,
        # bit-identical); several threads = any order, exact only when 1/ppr is a power of two
        ref.ref_set_num_threads(threads)
        want = np.zeros(mesh.nverts)
        ref.ref_gyro_scatter(mesh.nverts, mesh.nelems, _i(mesh.elem2verts), cap, _i(slot_elem),
                             mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _i(fwd), C.c_long(fwd.shape[0]),
                             C.c_double(rmax), nrings, ppr, _d(want))
        if threads == 1 or ppr == 8:
            assert np.array_equal(got, want) and want.sum() > 0
        else:
            assert np.allclose(got, want, rtol=1e-12, atol=0)
    ref.ref_set_num_threads(nthreads)


def test_elliptical_push_equals_the_reference(ref):
    """ellipticalPush::setup / push (test/ellipticalPush.hpp:10-70), compiled unmodified: the major axis
    and angle (stored as float) and the pushed positions must be identical over several pushes (both
    sides call the same libm)."""
    from meshes import load_fixture
    mesh = load_fixture("xgc24k")
    rng = np.random.default_rng(6)
    cap = 4000
    slot_elem = rng.integers(0, mesh.nelems, cap).astype(np.int32)
    mask = (rng.random(cap) < 0.9).astype(np.uint8)
    c = mesh.coords[mesh.elem2verts[slot_elem]].mean(axis=1)
    X = np.zeros((3, cap))
    X[0], X[1] = c[:, 0], c[:, 1]
    h, k, d = 1.6447937, 0.02055826, 0.6                       # pseudoXGCm.cpp:470-473
    cls = mesh.class_id.astype(np.int32)
    assert cls.min() >= 1
    b0, p0 = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    b1, p1 = b0.copy(), p0.copy()
    orc.elliptical_setup(mask, X, b0, p0, h, k, d)
    fp = C.POINTER(C.c_float)
    ref.ref_elliptical_setup(cap, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(X), C.c_long(cap),
                             b1.ctypes.data_as(fp), p1.ctypes.data_as(fp), C.c_double(h), C.c_double(k), C.c_double(d))
    assert _same(b0, b1) and _same(p0, p1) and np.abs(b0[mask > 0]).max() > 0
    T0, T1 = np.zeros((3, cap)), np.zeros((3, cap))
    for it in range(4):
        orc.elliptical_push(slot_elem, mask, T0, b0, p0, cls, h, k, d, 0.5)
        ref.ref_elliptical_push(cap, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _d(T1), C.c_long(cap),
                                b1.ctypes.data_as(fp), p1.ctypes.data_as(fp), mesh.nelems, _i(cls), C.c_double(h),
                                C.c_double(k), C.c_double(d), C.c_double(0.5))
        assert _same(T0, T1) and _same(p0, p1)
    assert np.abs(T0[:, mask > 0]).max() > 0 and not T0[:, mask == 0].any()


def test_set_unsafe_procs_equals_the_reference(ref):
    """setUnsafeProcs (src/pumipic_ptcl_ops.hpp:33-53) compiled unmodified."""
    rng = np.random.default_rng(8)
    cap, ne = 3000, 500
    slot_elem = rng.integers(0, ne, cap).astype(np.int32)
    mask = (rng.random(cap) < 0.85).astype(np.uint8)
    elems = rng.integers(-1, ne, cap).astype(np.int32)
    safe = (rng.random(ne) < 0.6).astype(np.int32)
    owner = rng.integers(0, 4, ne).astype(np.int32)
    e0, p0 = orc.set_unsafe_procs(mask, elems, safe, owner, 2)
    e1, p1 = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    ref.ref_set_unsafe_procs(cap, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), _i(elems), ne, _i(safe),
                             _i(owner), 2, e1.ctypes.data_as(ip), p1.ctypes.data_as(ip))
    assert np.array_equal(e0, e1) and np.array_equal(p0, p1) and (p0 != 2).any()


def test_gather_helpers_equal_the_reference(ref):
    """interpolateTetVtx / findBCCoordsInTet (adjacency.hpp:772-809) and interpolate2d_field /
    interp2dVector / interpolate3d_field (utils.hpp:245-456) compiled unmodified.  Where the reference
    itself reads out of bounds the comparison stays inside its defined range: interpolateTetVtx with
    dof == 1 (for dof > 1 it indexes a 4-entry gather with d*dof+comp, adjacency.hpp:779-783) and
    interpolate3d_field with ny, nz >= 2 (utils.hpp:385-404)."""
    from meshes import kuhn_cube
    L = orc.lib()
    L.orc_interpolate_tet_vtx.restype = C.c_double
    L.orc_interpolate3d_field.restype = C.c_double
    ref.ref_interpolate2d_field.restype = C.c_double
    ref.ref_interpolate3d_field.restype = C.c_double
    rng = np.random.default_rng(9)
    mesh = kuhn_cube(4)
    om = orc.OracleMesh(mesh)
    field = rng.random(mesh.nverts)
    for k in range(600):
        e = int(rng.integers(0, mesh.nelems))
        w = rng.random(4); w /= w.sum()
        xyz = w @ mesh.coords[mesh.elem2verts[e]]
        bcc = np.zeros(4)
        assert ref.ref_find_bcc_in_tet(_d(mesh.coords), mesh.nverts, _i(mesh.elem2verts), mesh.nelems, _d(xyz), e,
                                       _d(bcc)) == 0
        mine = np.zeros(4)
        assert L.orc_find_barycentric_tet(_d(mesh.coords[mesh.elem2verts[e]]), _d(xyz), _d(mine)) == 1
        assert _same(bcc, mine)
        out = C.c_double()
        assert ref.ref_interpolate_tet_vtx(_i(mesh.elem2verts), mesh.nelems, _d(field), C.c_long(field.shape[0]), e,
                                           _d(bcc), 1, 0, C.byref(out)) == 0
        assert out.value == L.orc_interpolate_tet_vtx(om.h, _d(field), e, _d(bcc), 1, 0)
    # a point outside its element: the reference's OMEGA_H_CHECK fires
    far = mesh.coords[mesh.elem2verts[0]].mean(axis=0) + 5.0
    assert ref.ref_find_bcc_in_tet(_d(mesh.coords), mesh.nverts, _i(mesh.elem2verts), mesh.nelems, _d(far), 0,
                                   _d(np.zeros(4))) == -2
    # regular 2D grids, 1 and 3 components, with and without cylindrical symmetry, points off the grid too
    nx, nz = 9, 7
    gx0, gz0, dx, dz = 0.3, -1.0, 0.25, 0.4
    for ncomp in (1, 3):
        data = rng.random(nx * nz * ncomp)
        for k in range(400):
            pos = np.array([rng.uniform(-0.5, 3.0), rng.uniform(-1.0, 1.0), rng.uniform(-2.0, 2.5)])
            for cyl in (0, 1):
                for comp in range(ncomp):
                    a = orc.interpolate2d_field(data, gx0, gz0, dx, dz, nx, nz, pos, cyl, ncomp, comp)
                    b = ref.ref_interpolate2d_field(_d(data), C.c_long(data.shape[0]), C.c_double(gx0), C.c_double(gz0),
                                                    C.c_double(dx), C.c_double(dz), nx, nz, _d(pos), cyl, ncomp, comp)
                    assert _same(a, b), (pos, cyl, comp)
                if ncomp == 3:
                    fa, fb = np.zeros(3), np.zeros(3)
                    L.orc_interp2d_vector(_d(data), C.c_double(gx0), C.c_double(gz0), C.c_double(dx), C.c_double(dz),
                                          nx, nz, _d(pos), _d(fa), cyl)
                    ref.ref_interp2d_vector(_d(data), C.c_long(data.shape[0]), C.c_double(gx0), C.c_double(gz0),
                                            C.c_double(dx), C.c_double(dz), nx, nz, _d(pos), _d(fb), cyl)
                    assert _same(fa, fb)
    # 3D grids
    gx, gy, gz = np.linspace(0, 1, 6), np.linspace(-1, 1, 5), np.linspace(2, 3, 4)
    data = rng.random(6 * 5 * 4)
    for k in range(800):
        x, y, z = rng.uniform(-0.2, 1.2), rng.uniform(-1.3, 1.3), rng.uniform(1.8, 3.2)
        a = orc.interpolate3d_field(x, y, z, gx, gy, gz, data)
        b = ref.ref_interpolate3d_field(C.c_double(x), C.c_double(y), C.c_double(z), 6, 5, 4, _d(gx), _d(gy), _d(gz),
                                        _d(data))
        assert _same(a, b)


@pytest.mark.parametrize("meshname", ["kuhn4", "cube7k", "plate15", "xgc24k"])
def test_workload_generator_equals_the_reference(ref, meshname):
    """pumi-pic_b200/workloads.py (the synthetic inputs of the tests and of bench.py) against the
    reference's own generator, test/test_adj.cpp (setSourceElements :27-42, init2DInternal :68-125,
    init3DInternal :440-507, get_push_distance :543-547, push_ptcls :550-562), compiled unmodified:
    std::default_random_engine(512*512) drawn per SLOT, the fold into the simplex, the direction, the
    push distance and the push itself must give identical doubles."""
    import ptcl_init as pi
    from meshes import kuhn_cube, load_fixture, plate
    mesh = {"kuhn4": lambda: kuhn_cube(4), "plate15": lambda: plate(15)}.get(meshname, lambda: load_fixture(meshname))()
    dim = mesh.dim
    ref.ref_testadj_push_distance.restype = C.c_double
    # particles per element
    for nptcls in (mesh.nelems * 3 + 17, 5):
        ppe = np.zeros(mesh.nelems, np.int32)
        assert ref.ref_testadj_ppe(mesh.nelems, nptcls, ppe.ctypes.data_as(ip)) == nptcls
        assert np.array_equal(ppe, pi.even_ppe(mesh.nelems, nptcls))
    d_ref = ref.ref_testadj_push_distance(dim, mesh.nverts, _d(mesh.coords), mesh.nelems)
    assert d_ref == pi.push_distance(mesh)
    # a padded structure: some slots empty
    rng = np.random.default_rng(12)
    cap = 6000
    slot_elem = np.sort(rng.integers(0, mesh.nelems, cap)).astype(np.int32)
    mask = (rng.random(cap) < 0.9).astype(np.uint8)
    init = pi.init3d_internal if dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    x, xt, mo = np.zeros((3, cap)), np.full((3, cap), 9.0), np.zeros((3, cap))
    pids = np.full(cap, -1, np.int32)
    ref.ref_testadj_init_internal(dim, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts),
                                  _d(orc.OracleMesh(mesh).vol()), cap, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_long(cap),
                                  _d(x), _d(xt), pids.ctypes.data_as(ip), _d(mo))
    m = mask.astype(bool)
    assert np.array_equal(X[:, m], x[:, m]) and np.array_equal(D[:, m], mo[:, m])
    assert np.array_equal(xt[:, m], x[:, m]) and np.array_equal(pids[m], np.flatnonzero(m))
    # the push: tgt += distance * direction on masked slots (the arithmetic of the fused kernel's PUSH form 1)
    T = xt.copy()
    T[:, m] = xt[:, m] + d_ref * mo[:, m]
    ref.ref_testadj_push(cap, _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_long(cap), _d(x), _d(xt),
                         pids.ctypes.data_as(ip), _d(mo), C.c_double(d_ref))
    assert np.array_equal(T, xt)
    P = np.zeros((3, cap)); P[:, m] = x[:, m]
    orc.push_direction(mask, P, mo, d_ref)
    assert np.array_equal(P[:, m], xt[:, m])


@pytest.mark.parametrize("meshname,nptcls", [("xgc24k", 100000), ("xgc24k", 1000), ("plate15", 3000), ("plate15", 0)])
def test_xgcm_load_generator_equals_the_reference(ref, meshname, nptcls):
    """BASELINE configs[3] (pseudoXGCm): the particle load of test/pseudoXGCm.cpp -- setSourceElements
    :167-222 (std::normal_distribution on std::default_random_engine(1024*1024), integer sigma, the
    overshoot / shortfall rules) and setInitialPtclCoords :224-264 (two uniforms per SLOT of
    engine(512*512), folded) -- compiled unmodified, against workloads.xgc_source_elements /
    xgc_initial_coords: identical counts and identical doubles."""
    import ptcl_init as pi
    from meshes import load_fixture, plate
    mesh = plate(15) if meshname == "plate15" else load_fixture(meshname)
    rng = np.random.default_rng(3)
    cls = mesh.class_id.astype(np.int32) if meshname == "xgc24k" else rng.integers(100, 180, mesh.nelems).astype(np.int32)
    for self_rank, owners in ((0, np.zeros(mesh.nelems, np.int32)),
                              (2, rng.integers(0, 4, mesh.nelems).astype(np.int32)),
                              (7, np.zeros(mesh.nelems, np.int32))):          # rank 7 owns nothing
        for mdl_face in (141, 1 << 30, -5):
            ppe_r = np.full(mesh.nelems, -3, np.int32)
            tot_r = ref.ref_xgcm_source_elements(mesh.nelems, _i(cls), _i(owners), self_rank, mdl_face, nptcls,
                                                 ppe_r.ctypes.data_as(ip))
            ppe, tot = pi.xgc_source_elements(cls, owners, self_rank, mdl_face, nptcls)
            marked = (cls <= mdl_face) & (owners == self_rank)
            if not marked.any():
                assert tot_r == 0 and tot == 0          # the reference returns before it writes ppe
                continue
            assert tot_r == tot and np.array_equal(ppe_r, ppe)
            assert tot == nptcls and not ppe[~marked].any()
    ppe, tot = pi.xgc_source_elements(cls, np.zeros(mesh.nelems, np.int32), 0, 141 if meshname == "xgc24k" else 150, nptcls)
    # a padded structure over that load: rows of 32 slots, so most rows end in empty slots
    rows = -(-ppe // 32) * 32
    slot_elem = np.repeat(np.arange(mesh.nelems, dtype=np.int32), rows)
    first = np.concatenate(([0], np.cumsum(rows)[:-1]))
    mask = (np.arange(slot_elem.shape[0]) - np.repeat(first, rows) < np.repeat(ppe, rows)).astype(np.uint8)
    cap = mask.shape[0]
    assert int(mask.sum()) == tot
    x = np.full((3, cap), 9.0)
    ref.ref_xgcm_initial_coords(mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), cap, _i(slot_elem),
                                mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_long(cap), _d(x))
    X = pi.xgc_initial_coords(mesh, slot_elem, mask)
    m = mask.astype(bool)
    assert np.array_equal(X[:, m], x[:, m]) and (x[:, ~m] == 9.0).all()
    if tot:
        om = orc.OracleMesh(mesh)
        # every particle lies in its row element (what check_initial_parents would test)
        ids = np.full(cap, -1, np.int32)
        found, ids, st = om.search_mesh_2d(slot_elem, mask, X, ids, looplimit=5)
        assert found and np.array_equal(ids[m], slot_elem[m])


def test_bench_cpu_leg_engines_agree(ref):
    """bench.py's CPU legs: the reference-source engine (ref_bench_step: the reference's push_ptcls +
    search_mesh, OpenMP stand-ins, carried-over element ids aliased in place) and the oracle port
    produce the same positions and element ids step after step of the ping-pong loop."""
    import ptcl_init as pi
    from meshes import kuhn_cube
    mesh = kuhn_cube(6)
    om = orc.OracleMesh(mesh)
    n = 30000
    ppe = pi.even_ppe(mesh.nelems, n)
    slot_elem = np.repeat(np.arange(mesh.nelems, dtype=np.int32), ppe)[:n]
    mask = np.ones(n, np.uint8)
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    dist = 2.0 * pi.push_distance(mesh)
    off, val = om.side2elem_off(), om.side2elem()
    ref.ref_bench_create.restype = C.c_void_p
    h = C.c_void_p(ref.ref_bench_create(
        3, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides, _i(mesh.elem2sides),
        _i(mesh.side2verts), _i(off), _i(val),
        np.ascontiguousarray(om.exposed(), np.int8).ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()), n,
        _i(slot_elem), mask.ctypes.data_as(C.POINTER(C.c_ubyte))))
    A0, B0 = X.copy(), np.zeros_like(X)
    A1, B1 = X.copy(), np.zeros_like(X)
    ids0, ids1 = None, np.full(n, -1, np.int32)
    for it in range(4):
        sgn = dist if it % 2 == 0 else -dist
        np.copyto(B0, A0)
        orc.push_direction(mask, B0, D, sgn)
        f0, ids0, _, _, _ = om.search_mesh(slot_elem, mask, A0, B0, elem_ids=ids0)
        f1 = ref.ref_bench_step(h, _d(A1), _d(B1), _d(D), C.c_long(n), C.c_double(sgn), ids1.ctypes.data_as(ip),
                                int(it == 0))
        assert bool(f1) == f0 and np.array_equal(B0, B1) and np.array_equal(ids0, ids1)
        A0, B0 = B0, A0
        A1, B1 = B1, A1
    assert (ids0 == -1).any() and (ids0 >= 0).any()
    ref.ref_bench_destroy(h)


def test_constant_push_and_position_update_equal_the_reference(ref):
    """push (constant vector) and updatePtclPositions of test/pseudoPushAndSearch.cpp (:87-118,
    :142-154) compiled unmodified: the oracle's orc_push_constant / orc_update_positions match, including
    the reference's `+ ptclUnique_d[pid]` (a zero array) and the update running over every slot."""
    rng = np.random.default_rng(13)
    cap = 5000
    slot_elem = rng.integers(0, 100, cap).astype(np.int32)
    mask = (rng.random(cap) < 0.85).astype(np.uint8)
    X = rng.normal(0, 3, (3, cap))
    T0, T1 = rng.normal(0, 1, (3, cap)), None
    T1 = T0.copy()
    dist, dvec = 3.275, (0.3, -0.0, 1.0)
    orc.push_constant(mask, X, T0, dist, dvec)
    ub = C.POINTER(C.c_ubyte)
    ref.ref_push_constant(cap, _i(slot_elem), mask.ctypes.data_as(ub), _d(X), _d(T1), C.c_long(cap), C.c_double(dist),
                          C.c_double(dvec[0]), C.c_double(dvec[1]), C.c_double(dvec[2]))
    assert np.array_equal(T0, T1) and not np.array_equal(T0[:, mask > 0], X[:, mask > 0])
    X0, X1 = X.copy(), X.copy()
    orc.update_positions(X0, T0)
    ref.ref_update_positions(cap, _i(slot_elem), mask.ctypes.data_as(ub), _d(X1), _d(T1), C.c_long(cap))
    assert np.array_equal(X0, X1) and np.array_equal(T0, T1) and not T0.any()


def test_scs_geometry_of_the_reference_matches_the_measured_run(ref):
    """chooseChunkHeight / constructChunks / constructOffsets (SCS_buildFns.h:4-153) compiled unmodified:
    for the bench workload (10 M particles over the 998 250 tets of the Kuhn cube, C = 32, sigma = INT_MAX,
    V = 1024, PAD_EVENLY 10 %) the reference's geometry has exactly the capacity the product's device
    build reported in the committed bench line (profiles/r1f_bench_n1.json).  Plus the arithmetic of the
    three padding strategies on a small hand-checked case."""
    import json
    import ptcl_init as pi
    import scs_ref_layout as srl
    line = json.load(open(os.path.join(ROOT, "profiles", "r1f_bench_n1.json")))
    lay = srl.layout(pi.even_ppe(line["config"]["tets_per_gpu"], line["config"]["particles_per_gpu"]))
    assert lay["capacity"] == line["detail"]["capacity"] == 10998528
    assert lay["C"] == 32 and lay["nchunks"] == 31196 == lay["nslices"]
    # 70 rows, C = 4 (sigma = full sort): widths are the maxima of 4 sorted rows, then the padding
    ppe = np.array([0] * 10 + [1] * 20 + [5] * 30 + [9] * 8 + [40] * 2, np.int32)
    rng = np.random.default_rng(0)
    rng.shuffle(ppe)
    raw = np.sort(ppe).reshape(-1, 1)[: 68].reshape(17, 4).max(axis=1).tolist() + [40]
    for strat, pad in ((0, 0.1), (1, 0.1), (2, 0.1), (0, 0.0)):
        lay = srl.layout(ppe, max_c=4, V=8, shuffle_padding=pad, pad_strat=strat)
        w = np.asarray(raw)
        if pad > 0:
            if strat == 0:
                w = np.where(w > 0, w + int(w.sum() * pad / (w > 0).sum()), w)
            elif strat == 1:
                w = (w + w * pad).astype(np.int64)            # int += double: truncation
            else:
                cw2 = w.sum() / (1.0 / w[w > 0]).sum() * pad
                w = np.where(w > 0, (w + cw2 / np.maximum(w, 1)).astype(np.int64), w)
        assert lay["C"] == 4 and lay["nchunks"] == 18 and lay["chunk_widths"].tolist() == w.tolist(), strat
        nsl = sum(-(-int(x) // 8) for x in w)
        assert lay["nslices"] == nsl and lay["capacity"] == 4 * int(w.sum()) == lay["offsets"][-1]
        assert lay["num_empty"] == 10 + 2                       # empty elements + the 2 padding rows


def _up_csr(nents, e2k):
    """ask_up(k, dim): entity -> elements, ascending element ids."""
    ne, per = e2k.shape
    flat = e2k.ravel()
    order = np.argsort(flat, kind="stable")
    off = np.zeros(nents + 1, np.int32)
    np.cumsum(np.bincount(flat, minlength=nents), out=off[1:])
    return off, (order // per).astype(np.int32)


def _elem_to(full, k):
    """elements -> their entities of dimension k"""
    dim = full.dim
    if k == 0:
        return np.ascontiguousarray(full.ent2verts(dim), np.int32)
    if k == dim - 1:
        return np.ascontiguousarray(full.down(dim), np.int32)
    edges = full.down(2)[full.down(3)].reshape(full.nents(3), 12)      # every edge of a tet twice
    return np.ascontiguousarray(np.sort(edges, axis=1)[:, ::2], np.int32)


@pytest.mark.parametrize("case", ["plate2d", "kuhn3d"])
def test_picpart_setup_kernels_equal_the_reference(ref, case):
    """the set-up kernels of PICpart construction -- BFS / bfsBufferLayers / bfsSafeInward
    (src/pumipic_part_construct.cpp:387-468, every Input::bridge_dim), defineOwners :304-323,
    createGlobalNumbering :366-374, rankLidNumbering :376-385 -- compiled unmodified, against the
    product's host code (pp_host_picpart_tags_bridged, pp_host_picpart_build_bridged)."""
    import importlib
    P = importlib.import_module("pumi-pic_b200")
    if case == "plate2d":
        coords, elems = P.host_plate(12)
        dim, nranks = 2, 4
    else:
        coords, elems = P.host_kuhn_cube(5)
        dim, nranks = 3, 5
    full = P.HostMesh.from_elems(dim, coords, elems)
    ne = full.nents(dim)
    cx = coords[elems].mean(axis=1)
    rng = np.random.default_rng(8)
    slabs = np.minimum((cx[:, 0] / cx[:, 0].max() * nranks).astype(np.int32), nranks - 1)
    blocks = (np.minimum((cx[:, 0] / cx[:, 0].max() * 2).astype(np.int32), 1) * 2
              + np.minimum((cx[:, 1] / cx[:, 1].max() * 2).astype(np.int32), 1)).astype(np.int32) % nranks
    noisy = np.where(rng.random(ne) < 0.03, rng.integers(0, nranks, ne), slabs).astype(np.int32)
    FULL, BFS, MINIMUM, NONE = 0, 1, 2, 3
    combos = [(BFS, BFS, 3, 1), (BFS, BFS, 1, 2), (BFS, FULL, 2, 1), (BFS, FULL, 1, 3), (MINIMUM, MINIMUM, 3, 1),
              (FULL, BFS, 3, 2), (FULL, FULL, 3, 1), (NONE, NONE, 3, 1), (BFS, MINIMUM, 2, 2), (MINIMUM, BFS, 1, 1),
              (BFS, NONE, 2, 1), (FULL, MINIMUM, 0, 0), (BFS, BFS, 0, 0)]
    for bridge_dim in range(dim):
        e2b = _elem_to(full, bridge_dim)
        nb = full.nents(bridge_dim)
        off, val = _up_csr(nb, e2b)
        for owner in (slabs, blocks, noisy):
            if len(np.unique(owner)) < nranks:
                continue
            for bm, sm, bl, sl in combos:
                # the Input constructor's adjustments (pumipic_input.cpp:96-110), applied for the reference call
                rbm = MINIMUM if bm == NONE else bm
                rbl = 0 if rbm == MINIMUM else bl
                rsl = 0 if sm == MINIMUM else sl
                for rank in range(nranks):
                    safe_r = np.full(ne, -1, np.int32)
                    part_r = np.full(nranks, -1, np.int32)
                    ref.ref_picpart_tags(dim, bridge_dim, nb, _i(off), _i(val), ne, _i(owner), nranks, rank, rbm, sm,
                                         rbl, rsl, safe_r.ctypes.data_as(ip), part_r.ctypes.data_as(ip))
                    safe, part = P.host_picpart_tags_bridged(nb, e2b, owner, nranks, rank, bm, sm, bl, sl)
                    assert np.array_equal(safe != 0, safe_r != 0), (bridge_dim, bm, sm, bl, sl, rank)
                    assert np.array_equal(part != 0, part_r != 0), (bridge_dim, bm, sm, bl, sl, rank)
                    if bridge_dim == 0:          # the vertex-bridged entry point is the same function
                        s0, p0 = P.host_picpart_tags(dim, full.nents(0), elems, owner, nranks, rank, bm, sm, bl, sl)
                        assert np.array_equal(s0, safe) and np.array_equal(p0, part)
        # the PICpart record built with this bridge dimension carries the same safe zone and parts
        for rank in (0, nranks - 1):
            pic = P.Picpart.build(full, slabs, nranks, rank, BFS, BFS, 2, 1, bridge_dim=bridge_dim)
            safe, part = P.host_picpart_tags_bridged(nb, e2b, slabs, nranks, rank, BFS, BFS, 2, 1)
            info = pic.dim_info(dim)
            l2g = info["ent_l2g"]
            assert np.array_equal(l2g, np.flatnonzero(part[slabs] != 0))
            assert np.array_equal(pic.mesh().tag(dim, "safe") != 0, safe[l2g] != 0)
    # ownership and numbering of every dimension (a fully buffered PICpart keeps the full mesh's order)
    pic = P.Picpart.build(full, noisy, nranks, 0, FULL, FULL, 3, 1)
    m = pic.mesh()
    for k in range(dim + 1):
        n = full.nents(k)
        assert m.nents(k) == n
        if k < dim:
            off, val = _up_csr(n, _elem_to(full, k))
            own_r = np.full(n, -1, np.int32)
            ref.ref_define_owners(dim, k, n, _i(off), _i(val), ne, _i(noisy), nranks, own_r.ctypes.data_as(ip))
        else:
            own_r = noisy
        assert np.array_equal(m.tag(k, "ownership"), own_r)
        offs = np.zeros(nranks + 1, np.int32)
        gids = np.zeros(n, np.int64)
        lids = np.zeros(n, np.int32)
        ref.ref_global_numbering(n, _i(own_r), nranks, offs.ctypes.data_as(ip),
                                 gids.ctypes.data_as(C.POINTER(C.c_longlong)), lids.ctypes.data_as(ip))
        assert np.array_equal(m.tag(k, "gids"), gids) and np.array_equal(m.tag(k, "rank_lids"), lids)
        assert np.array_equal(pic.dim_info(k)["offset_ents_per_rank"], offs)


def test_cpn_ownership_equals_the_reference(ref, tmp_path):
    """`.cpn` partitions: the file reader (pumipic_input.cpp:61-85 builds the class-owner table) followed
    by setOwnerByClassification (part_construct.cpp:278-301, compiled unmodified) against
    pp_host_read_partition on the xgc/24k class ids."""
    import importlib
    from meshes import load_fixture
    P = importlib.import_module("pumi-pic_b200")
    mesh = load_fixture("xgc24k")
    cls = np.ascontiguousarray(mesh.class_id, np.int32)
    ids = np.unique(cls)
    rng = np.random.default_rng(5)
    nranks = 4
    owners_of = {int(c): int(rng.integers(0, nranks)) for c in ids}
    size = int(ids.max())
    cpn = tmp_path / "xgc.cpn"
    cpn.write_text("%d\n" % size + "".join("%d %d\n" % (c, o) for c, o in owners_of.items()))
    got = P.host_read_partition(cpn, mesh.nelems, cls)
    table = np.zeros(size + 1, np.int32)                      # host_owners(size+1), :73-76
    for c, o in owners_of.items():
        table[c] = o
    for self_rank in range(nranks):
        if not (got == self_rank).any():
            continue                                          # the reference asserts on a rank that owns nothing
        want = np.full(mesh.nelems, -1, np.int32)
        ref.ref_owner_by_classification(2, mesh.nelems, _i(cls), size + 1, _i(table), self_rank,
                                        want.ctypes.data_as(ip))
        assert np.array_equal(got, want)
