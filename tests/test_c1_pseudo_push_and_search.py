"""CPU-only: BASELINE configs[0] (test/pseudoPushAndSearch on cube/7k.osh) run through the oracle and,
iteration by iteration, through the reference's own source (oracle/_ref: its constant push, its
legacy 3D search_mesh and its updatePtclPositions compiled unmodified).  200 particles is the case the
reference registers (testing.cmake:106-108) and runs to completion; with the 100 000 particles
BASELINE.json quotes, the reference itself aborts in its second iteration (see below), so that run is
followed up to the abort.  The GPU side of the 200-particle case is in tests/test_zz_mirror_gpu.py."""
import ctypes as C
import os

import numpy as np
import pytest

import c1_case
import oracle_api as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpumipic_ref_primitives.so")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return np.ascontiguousarray(a, np.int32).ctypes.data_as(ip)


@pytest.mark.parametrize("num_ptcls", [200, 100000])
def test_c1_loop_oracle_and_reference_source(num_ptcls):
    mesh, ppe, centroid, dist, d, marked = c1_case.setup(num_ptcls)
    assert len(marked) == 260 and ppe.sum() == num_ptcls              # 260 exposed faces of class 156
    assert marked[-1] == 7254 and ppe[7254] == num_ptcls // 260 + num_ptcls % 260
    assert np.all(ppe[marked[:-1]] == num_ptcls // 260)
    om = orc.OracleMesh(mesh)
    ref = C.CDLL(REF_LIB) if os.path.exists(REF_LIB) else None
    off, val, doff, dval = om.side2elem_off(), om.side2elem(), om.dual_off(), om.dual()
    exposed = np.ascontiguousarray(om.exposed(), np.int8)
    # dense structure: slot i = particle i, grouped by element (any layout is legal)
    elem = np.repeat(np.arange(mesh.nelems, dtype=np.int32), ppe)
    pid = np.arange(num_ptcls, dtype=np.int32)
    X = np.ascontiguousarray(centroid[elem].T)
    history, aborted_at = [], None
    for it in range(1, c1_case.NUM_ITERATIONS + 1):
        n = elem.shape[0]
        if n == 0:
            break
        mask = np.ones(n, np.uint8)
        T = np.zeros((3, n))
        orc.push_constant(mask, X, T, dist, d)
        found, ids, xp, xf, st = om.search_mesh_legacy3d(elem, mask, X, T, looplimit=c1_case.MAX_LOOPS)
        if num_ptcls == 200:
            assert found and st.aborted == 0                          # assert(isFound), :207
        if st.aborted:
            # 100 000 particles seed all 260 marked elements; from some of them the legacy walk does
            # not converge within maxLoops (its fallback indexes the dual graph by face id,
            # adjacency.hpp:726), the particles are left in a wrong element and the next search hits
            # OMEGA_H_CHECK(false) (:619-627): the reference aborts here, and the oracle says so
            aborted_at = it
            if ref is not None:
                ids1 = np.full(n, -1, np.int32)
                xp1, xf1 = np.zeros(3 * n), np.full(n, -1, np.int32)
                ub = C.POINTER(C.c_ubyte)
                assert ref.ref_search_mesh_3d_variants(
                    0, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides,
                    _i(mesh.elem2sides), _i(mesh.side2verts), _i(off), _i(val),
                    exposed.ctypes.data_as(C.POINTER(C.c_byte)), _d(om.vol()), _i(doff), _i(dval), n, _i(elem),
                    mask.ctypes.data_as(ub), _d(X), _d(T), C.c_long(n), ids1.ctypes.data_as(ip), 1,
                    xp1.ctypes.data_as(dp), xf1.ctypes.data_as(ip), c1_case.MAX_LOOPS) == -2
            break
        if ref is not None:
            X1, T1 = X.copy(), np.zeros((3, n))
            ub = C.POINTER(C.c_ubyte)
            ref.ref_push_constant(n, _i(elem), mask.ctypes.data_as(ub), _d(X1), _d(T1), C.c_long(n),
                                  C.c_double(dist), C.c_double(d[0]), C.c_double(d[1]), C.c_double(d[2]))
            ids1 = np.full(n, -1, np.int32)
            xp1, xf1 = np.zeros(3 * n), np.full(n, -1, np.int32)
            r = ref.ref_search_mesh_3d_variants(
                0, mesh.nverts, _d(mesh.coords), mesh.nelems, _i(mesh.elem2verts), mesh.nsides,
                _i(mesh.elem2sides), _i(mesh.side2verts), _i(off), _i(val), exposed.ctypes.data_as(C.POINTER(C.c_byte)),
                _d(om.vol()), _i(doff), _i(dval), n, _i(elem), mask.ctypes.data_as(ub), _d(X1), _d(T1), C.c_long(n),
                ids1.ctypes.data_as(ip), 1, xp1.ctypes.data_as(dp), xf1.ctypes.data_as(ip), c1_case.MAX_LOOPS)
            assert r == int(found) and np.array_equal(T, T1) and np.array_equal(ids, ids1)
            assert np.array_equal(xf, xf1) and np.array_equal(xp, xp1.reshape(n, 3))
            ref.ref_update_positions(n, _i(elem), mask.ctypes.data_as(ub), _d(X1), _d(T1), C.c_long(n))
        orc.update_positions(X, T)
        if ref is not None:
            assert np.array_equal(X, X1) and not T.any() and not T1.any()
        keep = ids >= 0                                               # rebuild: new_element == -1 deletes
        if num_ptcls == 200:   # particles only leave through exposed faces, with a wall point on that face
            assert np.all(xf[~keep] >= 0) and np.all(exposed[xf[~keep]] == 1)
        history.append((it, n, int(keep.sum())))
        elem, pid, X = ids[keep].astype(np.int32), pid[keep], np.ascontiguousarray(X[:, keep])
    if num_ptcls == 200:
        # the registered case (testing.cmake:106-108): 200 particles, all in element 7254, one shared
        # path; the cube is 65.5 high, 3.275 per push: they leave together within the 30 iterations
        assert elem.shape[0] == 0 and aborted_at is None and 10 <= len(history) <= 21, history
        assert all(a[2] in (0, a[1]) for a in history) and history[0][1] == num_ptcls
    else:
        assert aborted_at == 2 and history[0][1] == num_ptcls and history[0][2] < num_ptcls
