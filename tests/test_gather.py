"""Gather (field -> particle interpolation): adjacency.hpp:772-809 and pumipic_utils.hpp:186-456.

The reference has no test for these helpers (they are called from GITRm's push), so the oracle is
pinned on what linear interpolation must reproduce: a field that is linear in space comes back as
its own value at the particle position.  The GPU test then compares the CUDA kernels with the
oracle bit for bit (tolerance 0), except for the cylindrical rotation, where cos / sin / atan2
differ by a few ulp between libm and CUDA (stated tolerance 1e-14 relative).
"""
import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from meshes import kuhn_cube, load_fixture


def _particles(mesh, n):
    slot_elem = (np.arange(n, dtype=np.int64) * mesh.nelems // n).astype(np.int32)
    mask = np.ones(n, np.uint8)
    mask[::17] = 0
    X, _ = pi.init3d_internal(mesh, slot_elem, mask)
    return slot_elem, mask, X


def _grid2():
    nx, nz = 23, 17
    x0, z0, dx, dz = 0.05, -0.4, 0.061, 0.083
    gx = x0 + dx * np.arange(nx); gz = z0 + dz * np.arange(nz)
    a = np.array([[1.5, -2.0, 0.3], [0.7, 0.2, -1.1], [-0.4, 0.9, 2.2]])   # comp c: a[c,0]*x + a[c,1]*z + a[c,2]
    data = np.zeros((nz, nx, 3))
    for c in range(3):
        data[:, :, c] = a[c, 0] * gx[None, :] + a[c, 1] * gz[:, None] + a[c, 2]
    return nx, nz, x0, z0, dx, dz, a, data.ravel()


def test_tet_vertex_field_reproduces_linear_fields():
    mesh = load_fixture("cube7k")
    om = orc.OracleMesh(mesh)
    slot_elem, mask, X = _particles(mesh, 5000)
    A = np.array([[0.3, -1.2, 2.0], [1.0, 0.5, -0.25], [-2.0, 0.1, 0.7]])
    b = np.array([0.5, -1.0, 3.0])
    field = (mesh.coords @ A.T + b)                      # [nverts, 3], vertex-major dof = 3
    out, bad = orc.gather_tet_field(om, mask, X, slot_elem, field.ravel(), 3)
    assert bad == 0
    m = mask.astype(bool)
    want = (A @ X[:, m]) + b[:, None]
    assert np.allclose(out[:, m], want, rtol=0, atol=1e-11)
    assert np.all(out[:, ~m] == 0)
    # dof = 1 is the case the reference indexes in bounds (adjacency.hpp:779-783)
    out1, bad1 = orc.gather_tet_field(om, mask, X, slot_elem, field[:, 1].copy(), 1)
    assert bad1 == 0 and np.array_equal(out1[0], out[1])
    # a particle outside its element is what the reference aborts on
    wrong = slot_elem.copy(); wrong[1] = (slot_elem[1] + mesh.nelems // 2) % mesh.nelems
    _, bad2 = orc.gather_tet_field(om, mask, X, wrong, field.ravel(), 3)
    assert bad2 == 1


def test_grid2d_reproduces_linear_fields_and_clamps():
    nx, nz, x0, z0, dx, dz, a, data = _grid2()
    rng = np.random.default_rng(7)
    for _ in range(200):
        p = np.array([x0 + rng.uniform(0, (nx - 1) * dx), rng.uniform(-1, 1), z0 + rng.uniform(0, (nz - 1) * dz)])
        for c in range(3):
            got = orc.interpolate2d_field(data, x0, z0, dx, dz, nx, nz, p, False, 3, c)
            assert abs(got - (a[c, 0] * p[0] + a[c, 1] * p[2] + a[c, 2])) < 1e-12
    # beyond both upper edges the corner value is returned (pumipic_utils.hpp:272-274)
    far = np.array([x0 + nx * dx, 0.0, z0 + nz * dz])
    assert orc.interpolate2d_field(data, x0, z0, dx, dz, nx, nz, far, False, 3, 1) == data[(nx * nz - 1) * 3 + 1]
    # cylindrical symmetry: the radial coordinate is sqrt(x^2 + y^2)
    p = np.array([0.3, 0.4, 0.1])
    got = orc.interpolate2d_field(data, x0, z0, dx, dz, nx, nz, p, True, 3, 0)
    assert abs(got - (a[0, 0] * 0.5 + a[0, 1] * 0.1 + a[0, 2])) < 1e-12
    # single-point table
    assert orc.interpolate2d_field(np.array([4.0, 5.0, 6.0]), 0, 0, 1, 1, 1, 1, p, False, 3, 2) == 6.0


def test_grid3d_reproduces_trilinear_fields():
    gx = np.linspace(-1, 2, 14); gy = np.linspace(0, 1, 9); gz = np.linspace(3, 5, 11)
    f = lambda x, y, z: 1.0 + 2 * x - 3 * y + 0.5 * z + 0.25 * x * y - 0.75 * y * z + x * z + 0.1 * x * y * z
    data = f(gx[None, None, :], gy[None, :, None], gz[:, None, None]).ravel()     # i + j*nx + k*nx*ny
    rng = np.random.default_rng(3)
    for _ in range(300):
        x, y, z = rng.uniform(-1, 2), rng.uniform(0, 1), rng.uniform(3, 5)
        assert abs(orc.interpolate3d_field(x, y, z, gx, gy, gz, data) - f(x, y, z)) < 1e-11
    # degenerate directions (ny == 1 / nz == 1) fall back to the lower-dimensional interpolant (:415-416)
    d2 = f(gx[None, :], 0.0, gz[:, None]).ravel()
    assert abs(orc.interpolate3d_field(0.3, 9.0, 4.2, gx, np.array([0.0]), gz, d2) - f(0.3, 0.0, 4.2)) < 1e-11
    d1 = f(gx, 0.0, 3.0)
    assert abs(orc.interpolate3d_field(0.3, 9.0, 9.0, gx, np.array([0.0]), np.array([3.0]), d1) - f(0.3, 0, 3.0)) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["scs", "csr", "dps"])
def test_gather_kernels_match_oracle(kind):
    from gpu_common import dev, make_gpu_mesh, make_ps, pp, torch
    P = pp(); t = torch()
    mesh = kuhn_cube(7)
    om = orc.OracleMesh(mesh)
    gm = make_gpu_mesh(mesh)
    k = {"scs": P.capi.PP_PS_SCS, "csr": P.capi.PP_PS_CSR, "dps": P.capi.PP_PS_DPS}[kind]
    ps = make_ps(k, pi.even_ppe(mesh.nelems, 40000))
    slot_elem, mask = ps.slot_elem_and_mask()
    X, D = pi.init3d_internal(mesh, slot_elem, mask)
    m = mask.astype(bool)
    rng = np.random.default_rng(11)
    # --- tet vertex field, dof 3 and 1, with a few particles flagged outside / without element
    field = rng.standard_normal((mesh.nverts, 3))
    ids = slot_elem.copy()
    live = np.nonzero(m)[0]
    ids[live[::50]] = -1
    ids[live[7]] = (slot_elem[live[7]] + mesh.nelems // 2) % mesh.nelems
    out_o, bad_o = orc.gather_tet_field(om, mask, X, ids, field.ravel(), 3)
    out_g, bad_g = P.gather_tet_field(gm, ps, dev(X), dev(ids), dev(field.ravel()), 3)
    assert bad_g == bad_o == 1
    assert np.array_equal(out_g.cpu().numpy(), out_o)
    o1, _ = orc.gather_tet_field(om, mask, X, ids, field[:, 2].copy(), 1)
    g1, _ = P.gather_tet_field(gm, ps, dev(X), dev(ids), dev(field[:, 2].copy()), 1)
    assert np.array_equal(g1.cpu().numpy(), o1)
    # --- 2D grid, positions spill over every edge of the table so all four branches run
    nx, nz, x0, z0, dx, dz, a, data = _grid2()
    data = data + 0.01 * rng.standard_normal(data.shape)
    Xg = X.copy()
    Xg[0] = x0 - 0.2 + (nx * dx + 0.4) * rng.random(X.shape[1])
    Xg[1] = rng.uniform(-0.5, 0.5, X.shape[1])
    Xg[2] = z0 - 0.2 + (nz * dz + 0.4) * rng.random(X.shape[1])
    for cyl in (False, True):
        for c in range(3):
            want = np.zeros(X.shape[1])
            for s in np.nonzero(m)[0][::9]:
                want[s] = orc.interpolate2d_field(data, x0, z0, dx, dz, nx, nz, Xg[:, s].copy(), cyl, 3, c)
            got = P.gather_grid2d(ps, dev(Xg), dev(data), x0, z0, dx, dz, nx, nz, cyl, 3, c).cpu().numpy()
            sel = np.nonzero(m)[0][::9]
            assert np.array_equal(got[sel], want[sel])
            assert np.all(got[~m[:got.shape[0]]] == 0) if (~m).any() else True
        vo = orc.gather_grid2d_vector(mask, Xg, data, x0, z0, dx, dz, nx, nz, cyl)
        vg = P.gather_grid2d_vector(ps, dev(Xg), dev(data), x0, z0, dx, dz, nx, nz, cyl).cpu().numpy()
        if cyl:   # cos / sin / atan2: stated tolerance
            assert np.allclose(vg, vo, rtol=1e-14, atol=1e-14)
            assert np.array_equal(vg[2], vo[2])
        else:
            assert np.array_equal(vg, vo)
    # --- 3D grid
    gx = np.linspace(-0.1, 1.1, 19); gy = np.linspace(-0.1, 1.1, 13); gz = np.linspace(-0.1, 1.1, 16)
    d3 = rng.standard_normal(19 * 13 * 16)
    Xs = X + 0.3 * (rng.random(X.shape) - 0.5)          # some points leave the table: indices clamp
    o3 = orc.gather_grid3d(mask, Xs, d3, gx, gy, gz)
    g3 = P.gather_grid3d(ps, dev(Xs), dev(d3), dev(gx), dev(gy), dev(gz)).cpu().numpy()
    assert np.array_equal(g3, o3)
    gy1 = np.array([0.25])
    d31 = rng.standard_normal(19 * 16)
    o31 = orc.gather_grid3d(mask, Xs, d31, gx, gy1, gz)
    g31 = P.gather_grid3d(ps, dev(Xs), dev(d31), dev(gx), dev(gy1), dev(gz)).cpu().numpy()
    assert np.array_equal(g31, o31)
    # --- gather feeding the Boris push (pumipic_push.hpp:26-71): E from the tet field, B from the grid
    E = out_g
    B = P.gather_grid2d_vector(ps, dev(Xg), dev(data), x0, z0, dx, dz, nx, nz, False)
    pos = dev(X); prev = dev(X.copy()); vel = dev(D)
    P.push_boris(pos, prev, vel, E, B, 1e-9)
    po, pr, ve = X.copy(), X.copy(), D.copy()
    orc.push_boris(po, pr, ve, out_o, vo if False else orc.gather_grid2d_vector(mask, Xg, data, x0, z0, dx, dz, nx, nz, False), 1e-9)
    assert np.array_equal(pos.cpu().numpy(), po) and np.array_equal(vel.cpu().numpy(), ve)
