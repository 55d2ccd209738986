"""GPU tests of the particle structures, following particle_structs/test/{buildSCSTest.cpp,
test_structure.cpp,test_rebuild.cpp}: invariants keyed by particle id (slot numbering and row
order are backend-dependent in the reference)."""
import numpy as np
import pytest

from gpu_common import dev, pp, torch

pytestmark = pytest.mark.gpu

TYPES = [(np.int32, 1), (np.float64, 3), (np.int32, 1)]      # test particle: id, vec3, int


def _kinds():
    P = pp()
    return {
        "scs_c32": dict(kind=P.capi.PP_PS_SCS, team_size=32, sigma=0x7fffffff, V=1024),
        "scs_s1_v10": dict(kind=P.capi.PP_PS_SCS, team_size=32, sigma=1, V=10),
        "scs_c4_v2": dict(kind=P.capi.PP_PS_SCS, team_size=4, sigma=0x7fffffff, V=2),
        "scs_s7": dict(kind=P.capi.PP_PS_SCS, team_size=8, sigma=7, V=3),
        "scs_padprop": dict(kind=P.capi.PP_PS_SCS, team_size=32, config={"padding_strat": 1}),
        "scs_padinv": dict(kind=P.capi.PP_PS_SCS, team_size=32, config={"padding_strat": 2}),
        "csr": dict(kind=P.capi.PP_PS_CSR),
        "cabm": dict(kind=P.capi.PP_PS_CABM),
        "dps": dict(kind=P.capi.PP_PS_DPS),
    }


KINDS = ["scs_c32", "scs_s1_v10", "scs_c4_v2", "scs_s7", "scs_padprop", "scs_padinv", "csr", "cabm", "dps"]
SIZES = [(5, 25), (50, 1000), (2500, 100000), (1, 40), (300, 0)]


def _ppe(ne, np_, seed=3):
    rng = np.random.default_rng(seed)
    if np_ == 0:
        return np.zeros(ne, np.int32)
    w = rng.random(ne) ** 3          # skewed, with empty elements
    w[rng.random(ne) < 0.2] = 0
    if w.sum() == 0:
        w[0] = 1
    ppe = np.floor(w / w.sum() * np_).astype(np.int32)
    ppe[np.argmax(w)] += np_ - ppe.sum()
    return ppe


def _make(kindname, ne, np_, with_data=False, seed=3):
    P = pp()
    ppe = _ppe(ne, np_, seed)
    kw = dict(_kinds()[kindname])
    kind = kw.pop("kind")
    pel = info = None
    if with_data:
        pel = np.repeat(np.arange(ne, dtype=np.int32), ppe)
        rng = np.random.default_rng(seed + 1)
        rng.shuffle(pel)
        ids = np.arange(np_, dtype=np.int32)
        info = [ids.reshape(1, -1), rng.random((3, np_)), (ids * 7).reshape(1, -1)]
    ps = P.ParticleStructure(kind, TYPES, ppe, particle_elements=pel, particle_info=info, **kw)
    return ps, ppe, pel, info


@pytest.mark.parametrize("kindname", KINDS)
@pytest.mark.parametrize("ne,np_", SIZES)
def test_build_counts_and_layout(kindname, ne, np_):
    """buildSCSTest.cpp:47-129 / test_structure.cpp:43-113: sizes, per-element counts, mask."""
    ps, ppe, _, _ = _make(kindname, ne, np_)
    assert ps.nelems == ne and ps.nptcls == np_
    assert ps.capacity >= np_
    slot_elem, mask = ps.slot_elem_and_mask()
    assert mask.sum() == np_
    m = mask.astype(bool)
    assert np.array_equal(np.bincount(slot_elem[m], minlength=ne)[:ne], ppe)
    lay = ps.layout()
    if lay.kind in (0, 2) and np_ > 0:           # Sell-C-sigma geometry (SellCSigma.h:541-551)
        t = torch()
        from importlib import import_module
        api = import_module("pumi-pic_b200").api
        off = api._tensor_from_ptr(lay.offsets, (lay.nslices + 1,), t.int32, ps).cpu().numpy()
        s2c = api._tensor_from_ptr(lay.slice_to_chunk, (lay.nslices,), t.int32, ps).cpu().numpy()
        r2e = api._tensor_from_ptr(lay.row_to_element, (lay.nrows,), t.int32, ps).cpu().numpy()
        e2r = api._tensor_from_ptr(lay.element_to_row, (lay.nrows,), t.int32, ps).cpu().numpy()
        assert off[0] == 0 and off[-1] == ps.capacity and ps.numrows == lay.nchunks * lay.C
        assert np.all(np.diff(off) % lay.C == 0) and np.all(np.diff(off) // lay.C <= lay.V)
        assert np.all(np.diff(s2c) >= 0)
        assert np.array_equal(r2e[e2r[:ne]], np.arange(ne)) and np.array_equal(np.sort(r2e), np.arange(lay.nrows))
        # slot -> element through the slice geometry must equal the library's own map
        want = np.empty(ps.capacity, np.int32)
        for S in range(lay.nslices):
            n = off[S + 1] - off[S]
            rows = s2c[S] * lay.C + (np.arange(n) % lay.C)
            want[off[S]:off[S + 1]] = r2e[rows]
        assert np.array_equal(want, slot_elem)
        # sigma sort: ascending particle count inside each window (SCS_sort.h CUDA branch)
        sigma = min(_kinds()[kindname].get("sigma", 0x7fffffff), ne)
        counts = ppe[r2e[:ne]]
        if sigma > 1:
            for w0 in range(0, ne, sigma):
                assert np.all(np.diff(counts[w0:w0 + sigma]) >= 0)
                assert set(r2e[:ne][w0:w0 + sigma]) == set(range(w0, min(w0 + sigma, ne)))
        else:
            assert np.array_equal(r2e[:ne], np.arange(ne))


@pytest.mark.parametrize("kindname", ["scs_c32", "scs_c4_v2", "dps"])
def test_build_with_particle_data(kindname):
    """initSCSData / fillAoSoA: every given particle lands in its element with its data intact."""
    ne, np_ = 300, 20000
    ps, ppe, pel, info = _make(kindname, ne, np_, with_data=True)
    slot_elem, mask = ps.slot_elem_and_mask()
    m = mask.astype(bool)
    ids = ps.get(0).cpu().numpy()[0, :ps.capacity][m]
    assert np.array_equal(np.sort(ids), np.arange(np_))
    assert np.array_equal(slot_elem[m], pel[ids])
    assert np.array_equal(ps.get(1).cpu().numpy()[:, :ps.capacity][:, m], info[1][:, ids])
    assert np.array_equal(ps.get(2).cpu().numpy()[0, :ps.capacity][m], ids * 7)


def _state(ps):
    slot_elem, mask = ps.slot_elem_and_mask()
    return slot_elem, mask.astype(bool)


def _set_ids(ps):
    """pID(p) = p for every slot (the reference tests do this inside the lambda)."""
    t = torch()
    ids = ps.get(0)
    ids[0, :] = t.arange(ids.shape[1], dtype=t.int32, device="cuda")
    v = ps.get(1)
    v[:] = (t.arange(v.shape[1], dtype=t.float64, device="cuda") * 0.5)[None, :]


def _check(ps, new_element, cap0, removed=None, new_elems=None, np_expected=None):
    slot_elem, m = _state(ps)
    ids = ps.get(0).cpu().numpy()[0, :ps.capacity][m]
    vec = ps.get(1).cpu().numpy()[:, :ps.capacity][:, m]
    assert ps.nptcls == np_expected == m.sum()
    old = ids < cap0
    assert np.array_equal(slot_elem[m][old], new_element[ids[old]])      # destination check
    assert np.all(vec[:, old] == ids[old] * 0.5)                          # payload moved with it
    if removed is not None:
        assert not np.any(removed[ids[old]])
    if new_elems is not None:
        newi = ids[~old] - cap0
        assert np.array_equal(np.sort(newi), np.arange(len(new_elems)))
        assert np.array_equal(slot_elem[m][~old], new_elems[newi])
    assert len(np.unique(ids)) == len(ids)


@pytest.fixture
def _default_rebuild_modes():
    yield
    pp().lib().pp_ps_set_staged_rebuild(2)
    pp().lib().pp_ps_set_rebuild_tuning(0, 14)
    pp().lib().pp_ps_set_rank_sort_threshold(128)
    pp().lib().pp_ps_set_shuffling(1)
    pp().lib().pp_ps_set_rebuild_split_rows(1)


# move = how records travel in a re-layout: device-side layout + single-pass gather (default for
# Sell-C-sigma with sparse rows), staged records + atomics (the other sparse cases),
# staged records + sort-derived ranks (crowded rows), direct scatter (A/B fallback)
@pytest.mark.parametrize("move", ["gather", "fast_staged", "staged", "ranks", "direct"])
@pytest.mark.parametrize("kindname", KINDS)
@pytest.mark.parametrize("ne,np_", [(5, 25), (50, 1000), (2500, 100000)])
def test_rebuild_scenarios(kindname, ne, np_, move, _default_rebuild_modes):
    """test_rebuild.cpp: no change, new elements, added, deleted, added+deleted."""
    t = torch()
    pp().lib().pp_ps_set_staged_rebuild({"direct": 0, "gather": 2, "fast_staged": 2}.get(move, 1))
    pp().lib().pp_ps_set_rebuild_tuning(0, 0 if move == "fast_staged" else 1 << 20)
    pp().lib().pp_ps_set_rank_sort_threshold(1 if move == "ranks" else 1 << 30)
    ps, ppe, _, _ = _make(kindname, ne, np_)

    def setup():
        _set_ids(ps)
        slot_elem, m = _state(ps)
        cap = ps.capacity
        slots = np.arange(cap)
        return slot_elem, m, cap, slots

    # 1. rebuildNoChanges (:5-67): id sums per element are conserved
    slot_elem, m, cap, slots = setup()
    new_element = np.where(m, slot_elem, -1).astype(np.int32)
    sums = np.bincount(slot_elem[m], weights=slots[m], minlength=ne)
    ps.rebuild(dev(new_element))
    _check(ps, new_element, cap, np_expected=np_)
    se2, m2 = _state(ps)
    ids2 = ps.get(0).cpu().numpy()[0, :ps.capacity][m2]
    assert np.array_equal(np.bincount(se2[m2], weights=ids2, minlength=ne), sums)

    # 2. rebuildNewElems (:70-129): (3e + p) % ne
    slot_elem, m, cap, slots = setup()
    new_element = np.where(m, (slot_elem * 3 + slots) % ne, -1).astype(np.int32)
    ps.rebuild(dev(new_element))
    _check(ps, new_element, cap, np_expected=np_)

    # 3. rebuildNewPtcls (:132-205): cap/2 new particles
    slot_elem, m, cap, slots = setup()
    new_element = np.where(m, (slot_elem * 3 + slots + 2) % ne, -1).astype(np.int32)
    nnp = cap // 2
    new_elems = (np.arange(nnp) % ne).astype(np.int32)
    info = [dev((np.arange(nnp) + cap).astype(np.int32).reshape(1, -1)),
            t.zeros((3, nnp), dtype=t.float64, device="cuda"),
            t.zeros((1, nnp), dtype=t.int32, device="cuda")]
    ps.rebuild(dev(new_element), dev(new_elems), info)
    _check(ps, new_element, cap, new_elems=new_elems, np_expected=np_ + nnp)
    np_now = np_ + nnp

    # 4. rebuildPtclsDestroyed (:208-262): every 7th slot removed
    slot_elem, m, cap, slots = setup()
    removed = m & (slots % 7 == 0)
    new_element = np.where(m & ~removed, slot_elem, -1).astype(np.int32)
    ps.rebuild(dev(new_element))
    np_now -= int(removed.sum())
    _check(ps, new_element, cap, removed=removed, np_expected=np_now)

    # 5. rebuildNewAndDestroyed (:265-341)
    slot_elem, m, cap, slots = setup()
    removed = m & (slots % 7 == 0)
    new_element = np.where(m & ~removed, (3 * slot_elem + 7) % ne, -1).astype(np.int32)
    nnp = cap // 2
    new_elems = (np.arange(nnp) % ne).astype(np.int32)
    info = [dev((np.arange(nnp) + cap).astype(np.int32).reshape(1, -1)),
            t.zeros((3, nnp), dtype=t.float64, device="cuda"),
            t.zeros((1, nnp), dtype=t.int32, device="cuda")]
    ps.rebuild(dev(new_element), dev(new_elems), info)
    np_now += nnp - int(removed.sum())
    _check(ps, new_element, cap, removed=removed, new_elems=new_elems, np_expected=np_now)


@pytest.mark.parametrize("kindname", ["scs_c32", "csr", "dps"])
def test_rebuild_to_empty_and_refill(kindname):
    """SCS_rebuild.h:169-181: deleting everything keeps the structure usable."""
    t = torch()
    ne, np_ = 40, 900
    ps, ppe, _, _ = _make(kindname, ne, np_)
    ps.rebuild(dev(np.full(ps.capacity, -1, np.int32)))
    assert ps.nptcls == 0
    _, mask = ps.slot_elem_and_mask()
    assert mask.sum() == 0
    n = 333
    new_elems = (np.arange(n) * 5 % ne).astype(np.int32)
    info = [dev(np.arange(n, dtype=np.int32).reshape(1, -1)),
            t.ones((3, n), dtype=t.float64, device="cuda"), t.zeros((1, n), dtype=t.int32, device="cuda")]
    ps.rebuild(dev(np.full(max(ps.capacity, 1), -1, np.int32)), dev(new_elems), info)
    assert ps.nptcls == n
    slot_elem, mask = ps.slot_elem_and_mask()
    m = mask.astype(bool)
    ids = ps.get(0).cpu().numpy()[0, :ps.capacity][m]
    assert np.array_equal(slot_elem[m], new_elems[ids])


def test_rebuild_rejects_inactive_new_particles():
    t = torch()
    ps, _, _, _ = _make("scs_c32", 10, 100)
    slot_elem, mask = ps.slot_elem_and_mask()
    info = [t.zeros((1, 2), dtype=t.int32, device="cuda"), t.zeros((3, 2), dtype=t.float64, device="cuda"),
            t.zeros((1, 2), dtype=t.int32, device="cuda")]
    with pytest.raises(pp().PumipicError):
        ps.rebuild(dev(np.where(mask, slot_elem, -1).astype(np.int32)),
                   dev(np.array([3, -1], np.int32)), info)


@pytest.mark.parametrize("kindname", ["scs_c32", "scs_c4_v2", "csr"])
def test_search_runs_on_every_structure_kind(kindname):
    """The fused search walks SCS / CSR slot geometry exactly like the flat structure."""
    import oracle_api as orc
    import ptcl_init as pi
    from gpu_common import PARTICLE, make_gpu_mesh
    from meshes import kuhn_cube
    mesh = kuhn_cube(6)
    P = pp()
    kw = dict(_kinds()[kindname]); kind = kw.pop("kind")
    ps = P.ParticleStructure(kind, PARTICLE, pi.even_ppe(mesh.nelems, 30000), **kw)
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    stride = ps.get(0).shape[1]
    Xp = np.zeros((3, stride)); Xp[:, :ps.capacity] = X
    Dp = np.zeros((3, stride)); Dp[:, :ps.capacity] = D
    gm = make_gpu_mesh(mesh)
    t = torch()
    x = dev(Xp); d = dev(Dp); tg = t.zeros_like(x)
    ids = t.zeros(ps.capacity, dtype=t.int32, device="cuda")
    dist = pi.push_distance(mesh)
    r = P.push_direction_search(gm, ps, d, dist, x, tg, ids, elem_ids_empty=True, from_orig=True)
    T = np.zeros_like(X); m = mask.astype(bool)
    T[:, m] = X[:, m] + dist * D[:, m]
    found, ids_o, _, _, st = orc.OracleMesh(mesh).search_mesh(slot_elem, mask, X, T)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    assert (r.found, r.loops) == (int(found), st.loops)
    # and the structure follows the particles: rebuild with the new elements
    ps.get(2)[0, :ps.capacity] = t.arange(ps.capacity, dtype=t.int32, device="cuda")
    ps.rebuild(ids)
    se2, m2 = ps.slot_elem_and_mask()
    m2 = m2.astype(bool)
    pid = ps.get(2).cpu().numpy()[0, :ps.capacity][m2]
    assert np.array_equal(se2[m2], ids_o[pid]) and m2.sum() == (ids_o[m] >= 0).sum()


@pytest.mark.parametrize("kindname", ["scs_c32", "scs_s7", "scs_padprop", "cabm"])
def test_reshuffle_keeps_the_layout_when_movers_fit(kindname, _default_rebuild_modes):
    """SCS_rebuild.h:4-120: few movers, a few deletions and a few new particles fit into the row
    padding -> capacity, slot->element map and every non-moving particle's slot stay as they are;
    when the movers do not fit the full re-layout runs and gives the same particle sets."""
    t = torch()
    ne, np_ = 900, 60000
    ps, ppe, _, _ = _make(kindname, ne, np_)
    rng = np.random.default_rng(4)
    for rnd in range(3):
        _set_ids(ps)
        slot_elem, m = _state(ps)
        cap = ps.capacity
        slots = np.arange(cap)
        stay = slot_elem.copy()
        movers = m & (rng.random(cap) < 0.01)
        dels = m & ~movers & (rng.random(cap) < 0.01)
        new_element = np.where(m, stay, -1).astype(np.int32)
        # movers go to elements that certainly have room: the ones losing a particle this round
        targets = slot_elem[dels]
        if len(targets) == 0:
            targets = slot_elem[m][:1]
        new_element[movers] = rng.choice(targets, movers.sum())
        new_element[dels] = -1
        # each deletion frees one slot, but several movers may pick the same target: cap them
        room = np.bincount(slot_elem[dels], minlength=ne)
        take = np.zeros(ne, np.int64)
        for sidx in np.flatnonzero(movers):
            e = new_element[sidx]
            if take[e] < room[e] and e != slot_elem[sidx]:
                take[e] += 1
            else:
                new_element[sidx] = slot_elem[sidx]
        n_move = int((m & (new_element != slot_elem) & (new_element >= 0)).sum())
        lay0 = ps.layout()
        sig0 = (lay0.capacity, lay0.nchunks, lay0.nslices)
        ps.rebuild(dev(new_element))
        lay1 = ps.layout()
        assert (lay1.capacity, lay1.nchunks, lay1.nslices) == sig0        # in place
        _check(ps, new_element, cap, removed=dels, np_expected=int(m.sum() - dels.sum()))
        se2, m2 = _state(ps)
        assert np.array_equal(se2, slot_elem)                              # rows did not move
        ids = ps.get(0).cpu().numpy()[0, :cap]
        stayed = m & (new_element == slot_elem)
        assert np.array_equal(ids[stayed], slots[stayed]) and m2[stayed].all()   # stayers kept their slots
        assert n_move > 0
    # now far too many movers for the padding: falls back to the re-layout
    _set_ids(ps)
    slot_elem, m = _state(ps)
    cap = ps.capacity
    new_element = np.where(m, (slot_elem * 7 + np.arange(cap)) % ne, -1).astype(np.int32)
    ps.rebuild(dev(new_element))
    _check(ps, new_element, cap, np_expected=int(m.sum()))
    # shuffling switched off: same sets
    pp().lib().pp_ps_set_shuffling(0)
    _set_ids(ps)
    slot_elem, m = _state(ps)
    cap = ps.capacity
    new_element = np.where(m, slot_elem, -1).astype(np.int32)
    ps.rebuild(dev(new_element))
    _check(ps, new_element, cap, np_expected=int(m.sum()))


PIC_TYPES = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]   # x, xtgt, pid, dir


@pytest.mark.parametrize("move", ["gather", "fast_staged", "staged", "direct"])
@pytest.mark.parametrize("kindname", ["scs_c32", "csr", "dps", "scs_s7"])
def test_rebuild_remap_folds_update_positions(kindname, move, _default_rebuild_modes):
    """pp_ps_set_rebuild_remap([1, -1, 2, 3]): the rebuilt structure holds x = old xtgt, xtgt = 0 --
    updatePtclPositions (pseudoPushAndSearch.cpp:142-154) folded into the record move -- for kept and
    for added particles, on every structure kind and move mode; the remap is one-shot."""
    t = torch()
    P = pp()
    P.lib().pp_ps_set_staged_rebuild({"direct": 0, "gather": 2, "fast_staged": 2}.get(move, 1))
    P.lib().pp_ps_set_rebuild_tuning(0, 0 if move == "fast_staged" else 1 << 20)
    ne, np_ = 400, 9000
    ppe = _ppe(ne, np_, seed=5)
    kw = dict(_kinds()[kindname])
    kind = kw.pop("kind")
    pel = np.repeat(np.arange(ne, dtype=np.int32), ppe)
    rng = np.random.default_rng(17)
    X, T, D = rng.random((3, np_)), rng.random((3, np_)) + 2.0, rng.random((3, np_)) + 5.0
    ids = np.arange(np_, dtype=np.int32).reshape(1, -1)
    ps = P.ParticleStructure(kind, PIC_TYPES, ppe, particle_elements=pel, particle_info=[X, T, ids, D], **kw)
    for rnd in range(2):            # second round: no remap set any more
        se, m = ps.slot_elem_and_mask(); m = m.astype(bool)
        cap = ps.capacity
        pid = ps.get(2).cpu().numpy()[0, :cap]
        x0 = ps.get(0).cpu().numpy()[:, :cap].copy(); t0 = ps.get(1).cpu().numpy()[:, :cap].copy()
        d0 = ps.get(3).cpu().numpy()[:, :cap].copy()
        new_elem = np.where(m, (se * 5 + pid) % ne, -1).astype(np.int32)
        new_elem[m & (pid % 11 == 0)] = -1
        nnew = 700
        nel = rng.integers(0, ne, nnew).astype(np.int32)
        nX, nT, nD = rng.random((3, nnew)), rng.random((3, nnew)) + 2.0, rng.random((3, nnew)) + 5.0
        nid = (np.arange(nnew, dtype=np.int32) + 1000000 * (rnd + 1)).reshape(1, -1)
        if rnd == 0:
            ps.set_rebuild_remap([1, -1, 2, 3])
        ps.rebuild(dev(new_elem), dev(nel), [dev(nX), dev(nT), dev(nid), dev(nD)])
        want = {}
        for s_ in np.nonzero(m & (new_elem >= 0))[0]:
            want[int(pid[s_])] = (int(new_elem[s_]), x0[:, s_], t0[:, s_], d0[:, s_])
        for j in range(nnew):
            want[int(nid[0, j])] = (int(nel[j]), nX[:, j], nT[:, j], nD[:, j])
        se2, m2 = ps.slot_elem_and_mask(); m2 = m2.astype(bool)
        cap2 = ps.capacity
        pid2 = ps.get(2).cpu().numpy()[0, :cap2][m2]
        x2 = ps.get(0).cpu().numpy()[:, :cap2][:, m2]; t2 = ps.get(1).cpu().numpy()[:, :cap2][:, m2]
        d2 = ps.get(3).cpu().numpy()[:, :cap2][:, m2]
        assert sorted(pid2.tolist()) == sorted(want)
        for k, q in enumerate(pid2.tolist()):
            e, xo, to, do = want[q]
            assert se2[m2][k] == e and np.array_equal(d2[:, k], do)
            if rnd == 0:
                assert np.array_equal(x2[:, k], to) and not t2[:, k].any()
            else:
                assert np.array_equal(x2[:, k], xo) and np.array_equal(t2[:, k], to)
    with pytest.raises(P.PumipicError):
        ps.set_rebuild_remap([2, -1, 2, 3])          # member 2 has another type
    with pytest.raises(P.PumipicError):
        ps.set_rebuild_remap([1, 1, 2, 3])           # a member may feed one destination only


def _layout_arrays(ps):
    """row_to_element, offsets, slice_to_chunk of a Sell-C-sigma structure as numpy arrays"""
    t = torch()
    from importlib import import_module
    api = import_module("pumi-pic_b200").api
    lay = ps.layout()
    off = api._tensor_from_ptr(lay.offsets, (lay.nslices + 1,), t.int32, ps).cpu().numpy().copy()
    s2c = api._tensor_from_ptr(lay.slice_to_chunk, (lay.nslices,), t.int32, ps).cpu().numpy().copy()
    r2e = api._tensor_from_ptr(lay.row_to_element, (lay.nrows,), t.int32, ps).cpu().numpy().copy()
    return r2e, off, s2c


@pytest.mark.parametrize("ne,frac", [(40000, 0.1), (60000, 0.3), (300000, 0.01)])
def test_rebuild_of_a_mostly_empty_structure_sorts_only_the_occupied_rows(ne, frac, _default_rebuild_modes):
    """A PICpart that buffers the whole mesh holds particles in its own share of the rows only.  After a
    rebuild has seen that, the layout code sorts the non-empty rows alone and places the empty ones by a
    prefix sum (k_split_rows).  The layout must be the one the full stable sort gives -- row for row --
    and the particles must arrive (ids, payload, elements), also when the occupied set drifts, shrinks,
    and when it grows past the bound the previous rebuild suggested (fallback)."""
    t = torch()
    P = pp()
    lib = P.lib()
    lib.pp_ps_set_shuffling(0)
    rng = np.random.default_rng(ne)
    occupied = np.sort(rng.choice(ne, max(40, int(ne * frac)), replace=False))
    ppe = np.zeros(ne, np.int32)
    ppe[occupied] = rng.integers(1, 9, occupied.shape[0])

    def u01(e, rank, step, salt):
        """deterministic per (element, rank in its row, step): the order of a row's particles differs from run to run"""
        h = (e.astype(np.uint64) * np.uint64(2654435761) + rank.astype(np.uint64) * np.uint64(40503)
             + np.uint64(step * 97 + salt * 1000003)) * np.uint64(0x9E3779B97F4A7C15)
        return ((h >> np.uint64(40)) & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24)

    def run(split):
        lib.pp_ps_set_rebuild_split_rows(1 if split else 0)
        ps = P.ParticleStructure(P.capi.PP_PS_SCS, TYPES, ppe, team_size=32, sigma=0x7fffffff, V=1024)
        out = []
        occ = occupied
        for step in range(6):
            _set_ids(ps)
            slot_elem, m = _state(ps)
            cap = ps.capacity
            slots = np.arange(cap)
            # rank of a particle inside its element (by slot): decisions depend on (element, rank) only
            idx = slots[m][np.lexsort((slots[m], slot_elem[m]))]
            es = slot_elem[idx]
            first = np.r_[0, np.flatnonzero(np.diff(es)) + 1]
            rank = np.zeros(cap, np.int64)
            rank[idx] = np.arange(idx.shape[0]) - np.repeat(first, np.diff(np.r_[first, idx.shape[0]]))
            e64 = slot_elem.astype(np.int64)
            # most particles stay, some hop to another occupied element, a few leave; step 3 moves half of
            # them into arbitrary elements (more non-empty rows than the previous rebuild's bound allows:
            # the speculation fails and the rebuild falls back), step 4 removes most of them again
            dest = slot_elem.copy()
            hop = m & (u01(e64, rank, step, 1) < 0.3)
            dest[hop] = occ[(u01(e64, rank, step, 2)[hop] * occ.shape[0]).astype(np.int64) % occ.shape[0]]
            if step == 3:
                grow = m & (u01(e64, rank, step, 3) < 0.5)
                dest[grow] = (u01(e64, rank, step, 4)[grow] * ne).astype(np.int64) % ne
            gone = m & (u01(e64, rank, step, 5) < (0.6 if step == 4 else 0.03))
            new_element = np.where(m & ~gone, dest, -1).astype(np.int32)
            nnp = 50
            new_elems = occ[(np.arange(nnp) * 7919 + step) % occ.shape[0]].astype(np.int32)
            info = [dev((np.arange(nnp) + cap).astype(np.int32).reshape(1, -1)),
                    t.zeros((3, nnp), dtype=t.float64, device="cuda"),
                    t.zeros((1, nnp), dtype=t.int32, device="cuda")]
            ps.rebuild(dev(new_element), dev(new_elems), info)
            _check(ps, new_element, cap, removed=gone, new_elems=new_elems,
                   np_expected=int((m & ~gone).sum()) + nnp)
            se, mk = _state(ps)
            occ = np.unique(se[mk])
            out.append(_layout_arrays(ps) + (np.bincount(se[mk], minlength=ne),))
        return out

    a, b = run(True), run(False)
    for step, (x, y) in enumerate(zip(a, b)):
        for k, name in enumerate(("row_to_element", "offsets", "slice_to_chunk", "particles per element")):
            assert np.array_equal(x[k], y[k]), (step, name)
    lib.pp_ps_set_rebuild_split_rows(1)
