"""GPU parity of the 2D XGC-like path: elliptical push, search_mesh_2d, gyro ring map, gyro
scatter, setUnsafeProcs -- against the CPU oracle and the reference's known answers."""
import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from gpu_common import dev, make_gpu_mesh, pp, torch
from meshes import load_fixture, plate

pytestmark = pytest.mark.gpu

XGC_PARTICLE = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float32, 1), (np.float32, 1)]


def test_gyro_scatter_known_answer_on_gpu():
    """test/pseudoXGCm_scatter.cpp:116-178: one particle in element 0, 2 rings x 6 points, r=0.2,
    theta=15; center-only map => v3: 2.0, v2: 12.0, v8: 0.0, others 2/3; fwd == bkwd."""
    mesh = load_fixture("tri8_parDiag")
    P = pp()
    gm = make_gpu_mesh(mesh)
    rings, ppr = 2, 6
    fmap, st = P.gyro_ring_map(gm, 0.2, rings, ppr, 15.0)
    assert st.found
    om = orc.OracleMesh(mesh)
    found, fmap_o = om.gyro_ring_map(0.2, rings, ppr, 15.0)
    assert np.array_equal(fmap.cpu().numpy(), fmap_o)
    cmap = fmap.cpu().numpy().reshape(mesh.nverts, rings * ppr * 3).copy()
    for v in range(mesh.nverts):
        if v != 3:
            cmap[v, :] = 2                                  # modifyMappings :58-80
    ppe = np.zeros(mesh.nelems, np.int32); ppe[0] = 1
    for kind in (P.capi.PP_PS_SCS, P.capi.PP_PS_CSR, P.capi.PP_PS_DPS):
        ps = P.ParticleStructure(kind, XGC_PARTICLE, ppe, V=32)
        w = P.gyro_scatter(gm, ps, dev(cmap.ravel().astype(np.int32)), 0.2, rings, ppr).cpu().numpy()
        for v in range(mesh.nverts):
            want = {3: 2.0, 2: 12.0, 8: 0.0}.get(v, 2.0 / 3.0)
            assert abs(w[v] - want) <= 1e-10 * max(1.0, abs(want)), (v, w[v], want)   # are_close
        bk = P.gyro_scatter(gm, ps, dev(cmap.ravel().astype(np.int32)), 0.2, rings, ppr)
        sync = P.gyro_interleave(dev(w), bk).cpu().numpy()
        assert np.array_equal(sync[0::2], w) and np.array_equal(sync[1::2], bk.cpu().numpy())


@pytest.mark.parametrize("kindname", ["scs", "csr", "dps"])
def test_gyro_scatter_matches_oracle_xgc24k(kindname):
    """pseudoXGCm config (3 rings x 8 points, rmax 0.038): charges are dyadic => exact."""
    mesh = load_fixture("xgc24k")
    P = pp()
    gm = make_gpu_mesh(mesh)
    om = orc.OracleMesh(mesh)
    rings, ppr, rmax = 3, 8, 0.038
    fmap, st = P.gyro_ring_map(gm, rmax, rings, ppr, 0.0)
    found, fmap_o = om.gyro_ring_map(rmax, rings, ppr, 0.0)
    got = fmap.cpu().numpy()
    # cos/sin differ by <= 2 ulp between CUDA and glibc; a ring point on an edge could flip
    mism = (got.reshape(-1, 3) != fmap_o.reshape(-1, 3)).any(axis=1).mean()
    assert mism < 1e-4 and st.found == int(found)
    rng = np.random.default_rng(5)
    ppe = rng.poisson(4.0, mesh.nelems).astype(np.int32)
    ppe[mesh.class_id > 141] = 0
    kind = {"scs": P.capi.PP_PS_SCS, "csr": P.capi.PP_PS_CSR, "dps": P.capi.PP_PS_DPS}[kindname]
    ps = P.ParticleStructure(kind, XGC_PARTICLE, ppe)
    slot_elem, mask = ps.slot_elem_and_mask()
    w = P.gyro_scatter(gm, ps, dev(fmap_o), rmax, rings, ppr).cpu().numpy()
    w_o = om.gyro_scatter(slot_elem, mask, fmap_o, rmax, rings, ppr)
    assert np.array_equal(w, w_o)                    # exact: integer counts / 8
    # charge conservation: every particle puts 1 on rings 0 and 1 of its 3 vertices, and a ring's
    # value reaches the mesh once per mapped (point, vertex) pair divided by points-per-ring
    fm = fmap_o.reshape(mesh.nverts, rings, ppr, 3)
    acc = np.zeros((mesh.nverts, rings))
    np.add.at(acc, (mesh.elem2verts[slot_elem[mask.astype(bool)]].ravel(),), np.array([1.0, 1.0] + [0.0] * (rings - 2)))
    expect = (acc[:, :, None, None] * (fm >= 0)).sum() / ppr
    assert abs(w.sum() - expect) <= 1e-9 * max(1.0, expect)


def test_elliptical_push_search_rebuild_loop():
    """test/pseudoXGCm.cpp:504-534 on xgc/24k: push 0.5 deg, search_mesh_2d(maxLoops=200),
    updatePtclPositions, rebuild -- every step checked against the oracle."""
    mesh = load_fixture("xgc24k")
    P = pp()
    t = torch()
    gm = make_gpu_mesh(mesh)
    om = orc.OracleMesh(mesh)
    rng = np.random.default_rng(11)
    ppe = np.where(mesh.class_id <= 141, rng.integers(0, 6, mesh.nelems), 0).astype(np.int32)
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, XGC_PARTICLE, ppe, V=1024,
                             config={"extra_padding": 0.0})
    h, k, d = 1.6447937, 0.02055826, 0.6              # pseudoXGCm.cpp:470-473
    slot_elem, mask = ps.slot_elem_and_mask()
    X, _ = pi.init2d_internal(mesh, slot_elem, mask)
    cap = ps.capacity
    stride = ps.get(0).shape[1]
    ps.get(0)[:, :cap] = dev(X)
    ps.get(2)[0, :cap] = t.arange(cap, dtype=t.int32, device="cuda")
    P.elliptical_setup(ps, ps.get(0), ps.get(3), ps.get(4), h, k, d)
    b_o = np.zeros(cap, np.float32); phi_o = np.zeros(cap, np.float32)
    orc.elliptical_setup(mask, X, b_o, phi_o, h, k, d)
    b_g = ps.get(3).cpu().numpy()[0, :cap]; phi_g = ps.get(4).cpu().numpy()[0, :cap]
    m = mask.astype(bool)
    # atan2/sin differ by ulps between CUDA and glibc; members are stored as float
    assert np.allclose(b_g[m], b_o[m], rtol=2e-7, atol=0) and np.allclose(phi_g[m], phi_o[m], rtol=2e-7, atol=1e-7)
    Xo = X.copy()
    pid_of_slot = np.arange(cap)
    total_moved = 0
    for it in range(4):
        cap = ps.capacity
        slot_elem, mask = ps.slot_elem_and_mask(); m = mask.astype(bool)
        # oracle state is re-read from the device so that ulp-level push differences do not
        # accumulate into different walks: each step is compared on identical inputs
        Xo = ps.get(0).cpu().numpy()[:, :cap].copy()
        b_o = ps.get(3).cpu().numpy()[0, :cap].copy(); phi_o = ps.get(4).cpu().numpy()[0, :cap].copy()
        P.elliptical_push(gm, ps, ps.get(1), ps.get(3), ps.get(4), h, k, d, 0.5)
        To = np.zeros((3, cap))
        orc.elliptical_push(slot_elem, mask, To, b_o, phi_o, mesh.class_id, h, k, d, 0.5)
        Tg = ps.get(1).cpu().numpy()[:, :cap]
        assert np.allclose(Tg[:2, m], To[:2, m], rtol=1e-12, atol=1e-13)    # stated tolerance (cos/sin)
        ids = t.full((cap,), -1, dtype=t.int32, device="cuda")
        r = P.search_mesh(gm, ps, ps.get(0), ps.get(1), ids, variant=P.capi.PP_SEARCH_2D_LEGACY, looplimit=200)
        found, ids_o, st = om.search_mesh_2d(slot_elem, mask, Tg, np.full(cap, -1, np.int32), looplimit=200)
        assert r.found and found
        assert np.array_equal(ids.cpu().numpy(), ids_o)          # same inputs => bit-exact ids
        total_moved += int((ids_o[m] != slot_elem[m]).sum())
        P.update_positions(ps, ps.get(0), ps.get(1))
        pid_before = ps.get(2).cpu().numpy()[0, :cap][m]
        ps.rebuild(ids)
        se2, m2 = ps.slot_elem_and_mask(); m2 = m2.astype(bool)
        pid_after = ps.get(2).cpu().numpy()[0, :ps.capacity][m2]
        keep = ids_o[m] >= 0
        assert np.array_equal(np.sort(pid_after), np.sort(pid_before[keep]))
        lut = dict(zip(pid_before.tolist(), ids_o[m].tolist()))
        assert all(lut[p] == e for p, e in zip(pid_after.tolist(), se2[m2].tolist()))
    assert total_moved > 0


def test_set_unsafe_procs_matches_oracle():
    mesh = plate(16)
    P = pp()
    gm = make_gpu_mesh(mesh)
    rng = np.random.default_rng(2)
    owner = (np.arange(mesh.nelems) * 4 // mesh.nelems).astype(np.int32)
    safe = (rng.random(mesh.nelems) < 0.7).astype(np.int32)
    gm.set_picpart(safe, owner, 1)
    ps = P.ParticleStructure(P.capi.PP_PS_SCS, XGC_PARTICLE, pi.even_ppe(mesh.nelems, 5000))
    slot_elem, mask = ps.slot_elem_and_mask()
    elems = rng.integers(-1, mesh.nelems, ps.capacity).astype(np.int32)
    ne, npr = P.set_unsafe_procs(gm, ps, dev(elems))
    ne_o, npr_o = orc.set_unsafe_procs(mask, elems, safe, owner, 1)
    assert np.array_equal(ne.cpu().numpy(), ne_o) and np.array_equal(npr.cpu().numpy(), npr_o)
