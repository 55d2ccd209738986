"""CPU-only: the property checks of the reference's test/test_adj.cpp (:565-744, :828-948) applied to
the oracle's new-API search_mesh in 2D and 3D, vectorised in numpy and with a STRICT tolerance.

The reference's own `test_parent_elements` accidentally runs its containment check with tol = 1.0
(it passes `(…, true, tol)` into `(…, tol, debug)`, test_adj.cpp:577 vs adjacency.tpp:76), so it
pins very little; the 3D element ids have no golden vector in the reference either.
These checks are what pins the oracle's 3D walk on geometry alone:
  * a particle that stays ends in an element that contains its target (barycentric test, 1e-9);
  * a particle that leaves (BCC mode) has its target outside the mesh's bounding box -- on the
    convex cube / plate meshes used here that is equivalent to "outside the domain";
  * in intersection mode the wall point lies inside the reported exposed side and on the
    particle's path, the side belongs to the reported element, and particles without a wall hit
    are still inside the box (check_inside_bbox, :590-612).  NB the reference's 3D intersection
    walk tests a RAY (ray_intersects_triangle, adjacency.tpp:152-178, never bounds t from above):
    every moving particle is carried to the wall its ray meets, whatever its target; the 2D one
    tests the SEGMENT (line_edge_2d :204-218).  Both are restated as they are;
  * two consecutive pushes (testBCCSearch :746-781) keep every particle consistent.
"""
import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from meshes import kuhn_cube, load_fixture, plate

TOL = 1e-9


def _bcc(mesh, elems, pts):
    """barycentric coordinates (sum to 1) of pts[n, dim] in elements elems[n] -- numpy, independent of the oracle"""
    M = mesh.coords[mesh.elem2verts[elems]]                 # [n, dim+1, dim]
    A = np.concatenate([M, np.ones(M.shape[:2] + (1,))], axis=2).transpose(0, 2, 1)
    b = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
    return np.linalg.solve(A, b[:, :, None])[:, :, 0]


def _case(meshname, n, seed_shift=0):
    mesh = {"kuhn5": lambda: kuhn_cube(5), "plate15": lambda: plate(15)}.get(meshname, lambda: load_fixture(meshname))()
    om = orc.OracleMesh(mesh)
    slot_elem = ((np.arange(n, dtype=np.int64) * 7919 + seed_shift) % mesh.nelems).astype(np.int32)
    mask = np.ones(n, np.uint8)
    mask[::17] = 0
    init = pi.init3d_internal if mesh.dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    return mesh, om, slot_elem, mask, X, D


def _inside_box(mesh, P, tol):
    lo, hi = mesh.coords.min(axis=0), mesh.coords.max(axis=0)
    d = mesh.dim
    return np.all((P[:d].T >= lo - tol) & (P[:d].T <= hi + tol), axis=1)


@pytest.mark.parametrize("meshname,n", [("kuhn5", 6000), ("cube7k", 20000), ("plate15", 3000), ("tri8", 400)])
def test_bcc_search_two_pushes_properties(meshname, n):
    mesh, om, slot_elem, mask, X, D = _case(meshname, n)
    m = mask.astype(bool)
    dim = mesh.dim
    ids = None
    cur = X.copy()
    alive = m.copy()
    for push in range(2):                                    # testBCCSearch: push, search, push, search
        dist = (2.5 if push == 0 else -4.0) * pi.push_distance(mesh)
        T = cur.copy()
        T[:, m] = cur[:, m] + dist * D[:, m]
        found, ids, _, _, st = om.search_mesh(slot_elem, mask, cur, T, elem_ids=ids)
        assert found and st.not_in_elem == 0
        stay = alive & (ids >= 0)
        gone = alive & (ids < 0)
        assert stay.sum() > 0
        b = _bcc(mesh, ids[stay], T[:dim, stay].T)
        assert b.min() >= -TOL, "a particle ended outside its reported element"
        # the meshes are convex boxes: leaving the domain == target outside the bounding box
        assert not _inside_box(mesh, T[:, gone], -1e-12).any()
        assert _inside_box(mesh, T[:, stay], TOL).all()
        assert np.all(ids[~m] == -1)
        alive = stay
        cur = T
    assert (~alive & m).sum() > 0                             # some particles did leave


@pytest.mark.parametrize("meshname,n", [("kuhn5", 6000), ("cube7k", 20000), ("plate15", 3000)])
def test_intersection_search_properties(meshname, n):
    mesh, om, slot_elem, mask, X, D = _case(meshname, n, seed_shift=5)
    m = mask.astype(bool)
    dim = mesh.dim
    T = X.copy()
    T[:, m] = X[:, m] + 6.0 * pi.push_distance(mesh) * D[:, m]
    found, ids, faces, pts, st = om.search_mesh(slot_elem, mask, X, T, require_intersection=True)
    assert found and st.not_in_elem == 0
    hit = m & (faces >= 0)
    free = m & (faces < 0)
    if dim == 3:
        assert hit.sum() == m.sum()                           # a ray always reaches a wall
    else:
        assert hit.sum() > 20 and free.sum() > 20
    exposed = om.exposed()
    assert np.all(exposed[faces[hit]] == 1)
    assert np.all(ids[hit] >= 0), "a wall hit keeps its element (check_model_intersection :380-382)"
    # the wall side belongs to the reported element (test_wall_intersections :712-721)
    assert np.all((mesh.elem2sides[ids[hit]] == faces[hit][:, None]).any(axis=1))
    # the wall point lies inside the side ...
    sv = mesh.coords[mesh.side2verts[faces[hit]]]             # [k, dim, dim]
    xp = pts[hit]
    if dim == 3:
        e1, e2 = sv[:, 1] - sv[:, 0], sv[:, 2] - sv[:, 0]
        nrm = np.cross(e1, e2)
        scale = np.linalg.norm(nrm, axis=1)
        assert np.all(np.abs(np.einsum("ij,ij->i", nrm, xp - sv[:, 0])) <= 1e-9 * scale)
        # barycentric coordinates in the face (find_barycentric_tri_simple, :633-648)
        G = np.stack([np.einsum("ij,ij->i", e1, e1), np.einsum("ij,ij->i", e1, e2),
                      np.einsum("ij,ij->i", e2, e2)], axis=1)
        r = xp - sv[:, 0]
        r1, r2 = np.einsum("ij,ij->i", r, e1), np.einsum("ij,ij->i", r, e2)
        det = G[:, 0] * G[:, 2] - G[:, 1] ** 2
        u = (r1 * G[:, 2] - r2 * G[:, 1]) / det
        v = (r2 * G[:, 0] - r1 * G[:, 1]) / det
        assert np.all((u >= -1e-7) & (v >= -1e-7) & (u + v <= 1 + 1e-7))
    else:
        e = sv[:, 1] - sv[:, 0]
        r = xp - sv[:, 0]
        cross = r[:, 0] * e[:, 1] - r[:, 1] * e[:, 0]
        assert np.all(np.abs(cross) <= 1e-9 * np.linalg.norm(e, axis=1))
        t = np.einsum("ij,ij->i", r, e) / np.einsum("ij,ij->i", e, e)
        assert np.all((t >= -1e-7) & (t <= 1 + 1e-7))
    # ... and on the particle's path (:690-710): ahead of the origin, in 2D not beyond the target
    o, tg = X[:dim, hit].T, T[:dim, hit].T
    seg = tg - o
    lam = np.einsum("ij,ij->i", xp - o, seg) / np.einsum("ij,ij->i", seg, seg)
    assert np.all(lam >= -1e-9) and (dim == 3 or np.all(lam <= 1 + 1e-9))
    assert np.all(np.linalg.norm(o + lam[:, None] * seg - xp, axis=1) <= 1e-9 * np.maximum(1.0, lam))
    if dim == 3:
        assert lam.max() > 1.0                                # ... and often far beyond the target
        return
    # 2D: the target of a wall hit is outside the box, everything else stayed inside and sits in its element
    assert not _inside_box(mesh, T[:, hit], -1e-12).any()
    assert _inside_box(mesh, T[:, free], TOL).all()
    moved = free & (ids >= 0)
    b = _bcc(mesh, ids[moved], T[:dim, moved].T)
    assert b.min() >= -1e-7
    # both modes agree on who stays where, except for particles that graze a side
    fb, ids_b, _, _, _ = om.search_mesh(slot_elem, mask, X, T)
    assert fb and np.array_equal(ids_b < 0, (faces >= 0) | (~m))
    assert (ids_b[free] != ids[free]).mean() < 5e-3
