"""GPU parity of every barycentric-walk kernel variant on every structure kind.

pp_search_set_staged selects 0 = thread-per-slot kernel, 1 = block-staged kernel, 2 = the
Sell-C-sigma chunk walk (warp per chunk, row record in registers, per-warp hop queue).  All must
give results bit-identical to the CPU oracle (adjacency.tpp:642 search_mesh, BCC mode): element
ids, pushed positions and the counters (found, loops, not_in_elem, not_found, hops, active).
"""
import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from gpu_common import dev, make_gpu_mesh, make_ps, pp, torch
from meshes import kuhn_cube, load_fixture, plate

pytestmark = pytest.mark.gpu

KINDS = ["scs", "csr", "dps", "cabm"]


def _kind(name):
    c = pp().capi
    return {"scs": c.PP_PS_SCS, "csr": c.PP_PS_CSR, "dps": c.PP_PS_DPS, "cabm": c.PP_PS_CABM}[name]


@pytest.fixture(autouse=True)
def _restore_default_kernel():
    yield
    pp().lib().pp_search_set_staged(2)


def _mesh(name):
    return {"kuhn8": lambda: kuhn_cube(8), "plate20": lambda: plate(20),
            "kuhn3": lambda: kuhn_cube(3)}.get(name, lambda: load_fixture(name))()


def _uneven_ppe(nelems, nptcls, seed=7):
    """ragged rows: empty elements, a few crowded ones (exercises padding and wide chunks)"""
    rng = np.random.default_rng(seed)
    w = rng.exponential(1.0, nelems)
    w[rng.random(nelems) < 0.3] = 0.0
    ppe = np.floor(w / w.sum() * nptcls).astype(np.int32)
    ppe[rng.integers(0, nelems)] += 700        # one very wide row (> kQCap columns)
    return ppe


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("meshname,nptcls,ragged", [("cube7k", 60000, False), ("kuhn8", 40000, True),
                                                    ("xgc24k", 90000, True), ("plate20", 9000, False),
                                                    ("kuhn3", 37, False)])
def test_bcc_walk_variants_match_oracle(meshname, nptcls, ragged, kind, variant):
    mesh = _mesh(meshname)
    P = pp()
    P.lib().pp_search_set_staged(variant)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ppe = _uneven_ppe(mesh.nelems, nptcls) if ragged else pi.even_ppe(mesh.nelems, nptcls)
    ps = make_ps(_kind(kind), ppe)
    slot_elem, mask = ps.slot_elem_and_mask()
    init = pi.init3d_internal if mesh.dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    t = torch()
    m = mask.astype(bool)
    for mult in (1.0, 6.0):                       # short walks, then long ones (many hops)
        dist = mult * pi.push_distance(mesh)
        T = np.zeros_like(X)
        T[:, m] = X[:, m] + dist * D[:, m]
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, X, T)
        # separate search
        ids = t.full((ps.capacity,), -7, dtype=t.int32, device="cuda")
        r = P.search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True)
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert (r.found, r.loops, r.not_in_elem, r.not_found) == \
               (int(found), st.loops, st.not_in_elem, st.not_found)
        # fused push + search
        ids2 = t.full((ps.capacity,), -7, dtype=t.int32, device="cuda")
        tg = t.zeros(3, ps.capacity, dtype=t.float64, device="cuda")
        r2 = P.push_direction_search(gm, ps, dev(D), dist, dev(X), tg, ids2, elem_ids_empty=True,
                                     from_orig=True)
        assert np.array_equal(ids2.cpu().numpy(), ids_o)
        assert np.array_equal(tg.cpu().numpy()[:, m], T[:, m])
        assert (r2.found, r2.loops, r2.hops, r2.active) == (r.found, r.loops, r.hops, r.active)


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("kind", ["scs", "csr"])
def test_bcc_walk_variants_deletions_and_looplimit(kind, variant):
    mesh = load_fixture("cube7k")
    P = pp()
    P.lib().pp_search_set_staged(variant)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ps = make_ps(_kind(kind), pi.even_ppe(mesh.nelems, 30000))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    sl = np.flatnonzero(mask)
    bad = sl[::11]                                # not in their parent element
    far = (slot_elem[bad] + mesh.nelems // 2) % mesh.nelems
    X[:, bad] = mesh.coords[mesh.elem2verts[far]].mean(axis=1).T
    still = sl[5::17]                             # unmoved particles (finishUnmoved)
    T = X.copy()
    orc.push_direction(mask, T, D, 25.0)
    T[:, still] = X[:, still]
    t = torch()
    for limit in (0, 1, 3):
        ids = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
        r = P.search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True, looplimit=limit)
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, X, T, looplimit=limit)
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert (r.found, r.loops, r.not_in_elem, r.not_found) == \
               (int(found), st.loops, st.not_in_elem, st.not_found)
        assert r.not_in_elem > 0


def test_chunk_walk_after_rebuild_full_loop():
    """push+search -> update -> rebuild, five steps, Sell-C-sigma: particle sets by id must match
    the oracle's serial loop (slot numbering after a rebuild is backend-defined)."""
    mesh = kuhn_cube(6)
    P = pp()
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ppe = pi.even_ppe(mesh.nelems, 20000)
    ps = make_ps(P.capi.PP_PS_SCS, ppe)
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    m = mask.astype(bool)
    t = torch()
    x = ps.get(0); x.zero_(); x[:, :X.shape[1]] = dev(X)
    d = ps.get(3); d.zero_(); d[:, :X.shape[1]] = dev(D)
    pid = ps.get(2); pid.zero_(); pid[0, :X.shape[1]] = t.arange(X.shape[1], dtype=t.int32, device="cuda")
    # oracle state keyed by particle id
    o_pos = {int(s): X[:, s].copy() for s in np.flatnonzero(m)}
    o_dir = {int(s): D[:, s].copy() for s in np.flatnonzero(m)}
    o_elem = {int(s): int(slot_elem[s]) for s in np.flatnonzero(m)}
    dist = 1.7 * pi.push_distance(mesh)
    for step in range(5):
        cap = ps.capacity
        ids = t.zeros(cap, dtype=t.int32, device="cuda")
        x, tg, d, pid = ps.get(0), ps.get(1), ps.get(3), ps.get(2)
        r = P.push_direction_search(gm, ps, d, dist, x, tg, ids, elem_ids_empty=True, from_orig=True)
        assert r.not_in_elem == 0
        # oracle on its own flat arrays
        keys = sorted(o_pos)
        n = len(keys)
        Xo = np.array([o_pos[k] for k in keys]).T.copy() if n else np.zeros((3, 0))
        Do = np.array([o_dir[k] for k in keys]).T.copy() if n else np.zeros((3, 0))
        To = Xo + dist * Do
        se = np.array([o_elem[k] for k in keys], np.int32)
        found, ids_o, _, _, st = om.search_mesh(se, np.ones(n, np.uint8), Xo, To)
        # compare by particle id
        _, mk = ps.slot_elem_and_mask()
        live = np.flatnonzero(mk)
        got_pid = pid.cpu().numpy()[0, live]
        got = dict(zip(got_pid.tolist(), ids.cpu().numpy()[live].tolist()))
        want = dict(zip(keys, ids_o.tolist()))
        assert got == want
        tgh = tg.cpu().numpy()[:, live]
        by_pid = dict(zip(got_pid.tolist(), tgh.T))
        for j, k in enumerate(keys[:200]):
            assert np.array_equal(by_pid[k], To[:, j])
        P.update_positions(ps, x, tg)
        ps.rebuild(ids)
        for j, k in enumerate(keys):
            if ids_o[j] < 0:
                del o_pos[k], o_dir[k], o_elem[k]
            else:
                o_pos[k] = To[:, j]
                o_elem[k] = int(ids_o[j])
        assert ps.nptcls == len(o_pos)


@pytest.mark.parametrize("kind,nparts", [("scs", 5), ("scs", 1), ("scs", 64), ("csr", 4), ("dps", 3)])
def test_host_buffer_pipeline_matches_oracle(kind, nparts):
    """pp_push_direction_search_host: pinned host columns in, host results out, pieces overlapped."""
    mesh = load_fixture("cube7k")
    P = pp()
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ps = make_ps(_kind(kind), _uneven_ppe(mesh.nelems, 50000))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    m = mask.astype(bool)
    t = torch()
    dist = 2.5 * pi.push_distance(mesh)
    hx = t.as_tensor(X).pin_memory(); hd = t.as_tensor(D).pin_memory()
    ht = t.zeros_like(hx).pin_memory()
    hi = t.full((ps.capacity,), -9, dtype=t.int32).pin_memory()
    for rep in range(3):                           # later calls reuse the staging buffers; the third
        # one also the direction column already on the device (h_dir = NULL)
        ht.zero_(); hi.fill_(-9)
        r = P.push_direction_search_host(gm, ps, hx, hd if rep < 2 else None, ht, hi, dist, nparts=nparts)
        t.cuda.synchronize()
        T = np.zeros_like(X)
        T[:, m] = X[:, m] + dist * D[:, m]
        found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, X, T)
        assert np.array_equal(hi.numpy(), ids_o)
        assert np.array_equal(ht.numpy()[:, m], T[:, m])
        assert (r.found, r.loops, r.not_in_elem) == (int(found), st.loops, st.not_in_elem)
        assert r.active == int(m.sum())
    # a structure whose directions were never uploaded must be refused
    ps2 = make_ps(_kind(kind), _uneven_ppe(mesh.nelems, 1000))
    with pytest.raises(P.PumipicError):
        P.push_direction_search_host(gm, ps2, hx, None, ht, hi, dist, nparts=nparts)


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("kind", ["scs", "csr"])
@pytest.mark.parametrize("meshname,nptcls", [("xgc24k", 120000), ("plate20", 7000)])
def test_search_mesh_2d_fresh_ids_all_kernels(meshname, nptcls, kind, variant):
    """adjacency.hpp:1013 search_mesh_2d as test/pseudoXGCm.cpp:147-153 calls it: elem_ids is a fresh
    array of -1 (elem_ids_empty promises that), so particles start in their row element."""
    mesh = _mesh(meshname)
    P = pp()
    P.lib().pp_search_set_staged(variant)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ps = make_ps(_kind(kind), _uneven_ppe(mesh.nelems, nptcls, seed=3))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init2d_internal(mesh, slot_elem, mask)
    t = torch()
    for mult, limit in ((1.0, 200), (7.0, 200), (7.0, 2)):
        T = X.copy()
        orc.push_direction(mask, T, D, mult * pi.push_distance(mesh))
        start = np.full(ps.capacity, -1, np.int32)
        ids = dev(start)
        r = P.search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True,
                          variant=P.capi.PP_SEARCH_2D_LEGACY, looplimit=limit)
        found, ids_o, st = om.search_mesh_2d(slot_elem, mask, T, start, looplimit=limit)
        assert np.array_equal(ids.cpu().numpy(), ids_o)
        assert (r.found, r.loops, r.not_found) == (int(found), st.loops, st.not_found)


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("kind", ["scs", "csr", "dps"])
def test_search_mesh_2d_ids_empty_ignores_array_contents(kind, variant):
    """elem_ids_empty means "behave as if elem_ids.size()==0": the array's old contents (here
    garbage, including the -nelems sentinel) must not be read by any kernel variant on any
    structure kind; every particle starts in its row element."""
    mesh = _mesh("plate20")
    P = pp()
    P.lib().pp_search_set_staged(variant)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ps = make_ps(_kind(kind), _uneven_ppe(mesh.nelems, 7000, seed=4))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init2d_internal(mesh, slot_elem, mask)
    T = X.copy()
    orc.push_direction(mask, T, D, 3.0 * pi.push_distance(mesh))
    rng = np.random.default_rng(9)
    garbage = rng.integers(-mesh.nelems, mesh.nelems, ps.capacity).astype(np.int32)
    garbage[::7] = -mesh.nelems
    ids = dev(garbage)
    r = P.search_mesh(gm, ps, dev(X), dev(T), ids, elem_ids_empty=True,
                      variant=P.capi.PP_SEARCH_2D_LEGACY, looplimit=200)
    found, ids_o, st = om.search_mesh_2d(slot_elem, mask, T, np.full(ps.capacity, -1, np.int32), looplimit=200)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert (r.found, r.loops) == (int(found), st.loops)
    P.lib().pp_search_set_staged(2)


@pytest.mark.parametrize("V", [32, 100, 1024])
@pytest.mark.parametrize("meshname", ["kuhn8", "plate20"])
def test_chunk_walk_sliced_wide_rows(meshname, V):
    """The chunk walk's work unit is a vertical slice (at most V columns): a structure whose widest rows
    span many slices (pseudoXGCm's load puts ~1 M particles into one element) gives the oracle's ids,
    also through the host-buffer pipeline that cuts the structure at chunk boundaries."""
    mesh = _mesh(meshname)
    P = pp()
    P.lib().pp_search_set_staged(2)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    ppe = _uneven_ppe(mesh.nelems, 30000, seed=12)
    ppe[mesh.nelems // 3] += 4000
    ppe[mesh.nelems - 1] += 2500
    ps = make_ps(_kind("scs"), ppe, V=V)
    slot_elem, mask = ps.slot_elem_and_mask()
    init = pi.init3d_internal if mesh.dim == 3 else pi.init2d_internal
    X, D = init(mesh, slot_elem, mask)
    m = mask.astype(bool)
    t = torch()
    dist = 3.0 * pi.push_distance(mesh)
    T = np.zeros_like(X)
    T[:, m] = X[:, m] + dist * D[:, m]
    found, ids_o, _, _, st = om.search_mesh(slot_elem, mask, X, T)
    ids = t.full((ps.capacity,), -7, dtype=t.int32, device="cuda")
    tg = t.zeros(3, ps.capacity, dtype=t.float64, device="cuda")
    r = P.push_direction_search(gm, ps, dev(D), dist, dev(X), tg, ids, elem_ids_empty=True, from_orig=True)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert (r.found, r.loops, r.active) == (int(found), st.loops, int(m.sum()))
    if mesh.dim == 3:
        hx = t.as_tensor(X).pin_memory(); hd = t.as_tensor(D).pin_memory()
        ht = t.zeros_like(hx).pin_memory()
        hi = t.full((ps.capacity,), -9, dtype=t.int32).pin_memory()
        P.push_direction_search_host(gm, ps, hx, hd, ht, hi, dist, nparts=5)
        t.cuda.synchronize()
        assert np.array_equal(hi.numpy(), ids_o)
