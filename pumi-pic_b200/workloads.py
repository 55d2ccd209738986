"""Seeded particle initial conditions, following test/test_adj.cpp (numpy twins).

std_uniform() reproduces libstdc++'s std::default_random_engine (minstd_rand0) driving
std::uniform_real_distribution<double>(0,1): generate_canonical<double,53> draws two
31-bit values per double.
"""
import numpy as np

_M = np.uint64(2147483647)
_A = 16807


def _minstd_stream(seed, n):
    """x_1..x_n of minstd_rand0 started from `seed` (x_0)."""
    seed = seed % 2147483647
    if seed == 0:
        seed = 1
    B = 1 << 14
    mult = np.empty(B, np.uint64)          # A^(i+1) mod M
    a = 1
    for i in range(B):
        a = (a * _A) % 2147483647
        mult[i] = a
    jump = int(mult[-1])
    out = np.empty(n, np.uint64)
    x0 = seed
    for lo in range(0, n, B):
        k = min(B, n - lo)
        out[lo:lo + k] = (np.uint64(x0) * mult[:k]) % _M
        x0 = (x0 * jump) % 2147483647
    return out


def std_uniform(seed, n):
    raw = _minstd_stream(seed, 2 * n).astype(np.float64) - 1.0   # g() - min()
    R = 2147483646.0
    s = raw[0::2] + raw[1::2] * R
    r = s / (R * R)
    r[r >= 1.0] = np.nextafter(1.0, 0.0)
    return r


PARTICLE_SEED = 512 * 512


def init3d_internal(mesh, slot_elem, mask, seed=PARTICLE_SEED):
    """test/test_adj.cpp:440-507 init3DInternal: uniform point in the row's tet + unit direction."""
    cap = mask.shape[0]
    r = std_uniform(seed, 5 * cap).reshape(cap, 5)
    x, y, z, ang, rr = (r[:, i].copy() for i in range(5))
    f = x + y > 1
    x[f], y[f] = 1 - x[f], 1 - y[f]
    g = y + z > 1
    h = (~g) & (x + y + z > 1)
    tmp = z.copy()
    z[g] = 1 - x[g] - y[g]
    y[g] = 1 - tmp[g]
    zz = x + y + z - 1
    xx = 1 - y - tmp
    z[h] = zz[h]
    x[h] = xx[h]
    theta = ang * 2 * np.pi
    zdir = rr * 2 - 1
    se = np.where(mask.astype(bool), slot_elem, 0)     # padding rows carry element ids >= nelems
    V = mesh.coords[mesh.elem2verts[se]]               # [cap,4,3]
    a = 1 - x - y - z
    pos = (a[:, None] * V[:, 0] + x[:, None] * V[:, 1]) + y[:, None] * V[:, 2]
    pos = pos + z[:, None] * V[:, 3]
    X = np.zeros((3, cap))
    D = np.zeros((3, cap))
    m = mask.astype(bool)
    X[:, m] = pos[m].T
    D[0, m] = (np.sqrt(1 - zdir * zdir) * np.cos(theta))[m]
    D[1, m] = (np.sqrt(1 - zdir * zdir) * np.sin(theta))[m]
    D[2, m] = zdir[m]
    return X, D


def init2d_internal(mesh, slot_elem, mask, seed=PARTICLE_SEED):
    """test/test_adj.cpp:66-123 init2DInternal."""
    cap = mask.shape[0]
    r = std_uniform(seed, 3 * cap).reshape(cap, 3)
    x, y, ang = r[:, 0].copy(), r[:, 1].copy(), r[:, 2] * 2 * np.pi
    f = x + y > 1
    x[f], y[f] = 1 - x[f], 1 - y[f]
    se = np.where(mask.astype(bool), slot_elem, 0)
    V = mesh.coords[mesh.elem2verts[se]]               # [cap,3,2]
    pos = (V[:, 0] + x[:, None] * (V[:, 1] - V[:, 0])) + y[:, None] * (V[:, 2] - V[:, 0])
    X = np.zeros((3, cap))
    D = np.zeros((3, cap))
    m = mask.astype(bool)
    X[:2, m] = pos[m].T
    D[0, m] = np.cos(ang)[m]
    D[1, m] = np.sin(ang)[m]
    return X, D


def push_distance(mesh):
    """test/test_adj.cpp:541-547 get_push_distance."""
    ext = (mesh.coords.max(axis=0) - mesh.coords.min(axis=0)).max()
    if mesh.dim == 2:
        return ext / (3 * np.sqrt(mesh.nelems))
    return ext / (3 * mesh.nelems ** (1.0 / 3))


def even_ppe(nelems, nptcls):
    """test/test_adj.cpp:27-42 setSourceElements."""
    ppe = np.full(nelems, nptcls // nelems, np.int32)
    ppe[: nptcls % nelems] += 1
    return ppe
