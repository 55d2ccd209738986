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


ELEMENT_SEED = 1024 * 1024


def std_normal(seed, n):
    """The first n variates of libstdc++'s std::normal_distribution<double>(0, 1) driven by
    std::default_random_engine(seed): Marsaglia's polar method on pairs of
    generate_canonical<double,53> draws; an accepted pair (x, y) yields y * mult first and keeps
    x * mult for the next call (bits/random.tcc normal_distribution::operator())."""
    import math
    pairs = (n + 1) // 2
    k = int(pairs * 1.3) + 16                      # acceptance is pi / 4
    while True:
        u = std_uniform(seed, 2 * k).reshape(k, 2)
        x = 2.0 * u[:, 0] - 1.0
        y = 2.0 * u[:, 1] - 1.0
        r2 = x * x + y * y
        ok = ~((r2 > 1.0) | (r2 == 0.0))
        if int(ok.sum()) >= pairs:
            break
        k *= 2
    x, y, r2 = x[ok][:pairs], y[ok][:pairs], r2[ok][:pairs]
    # libm's log, element by element: numpy's vector log may differ from it in the last place
    mult = np.array([math.sqrt(-2 * math.log(v) / v) for v in r2.tolist()], np.float64).reshape(pairs)
    out = np.empty(2 * pairs)
    out[0::2] = y * mult
    out[1::2] = x * mult
    return out[:n]


def _std_round(v):
    """std::round: to nearest, halves away from zero (v - trunc(v) is exact)."""
    t = np.trunc(v)
    return t + np.sign(v) * (np.abs(v - t) >= 0.5)


def xgc_source_elements(class_id, owners, rank, mdl_face, nptcls, seed=ELEMENT_SEED):
    """test/pseudoXGCm.cpp:167-222 setSourceElements: particles per element, normal(mean = nptcls /
    marked, sigma = mean / 4 in INTEGER arithmetic) on every owned element whose class id is at most
    mdl_face, in element order until nptcls are placed; the overshoot is cut from the element that
    crossed the total, a shortfall goes to the last element touched.  Returns (ppe, total)."""
    nelems = class_id.shape[0]
    marked = (class_id <= mdl_face) & (owners == rank)
    ppe = np.zeros(nelems, np.int32)
    idx = np.flatnonzero(marked)
    if idx.shape[0] == 0 or nptcls <= 0:
        return ppe, 0
    nppe = nptcls // idx.shape[0]
    cnt = _std_round(std_normal(seed, idx.shape[0]) * float(nppe // 4) + float(nppe)).astype(np.int64)
    cnt[cnt < 0] = 0
    cum = np.cumsum(cnt)
    j = int(np.searchsorted(cum, nptcls, side="left"))        # the element whose draw reaches the total
    if j < idx.shape[0]:
        cnt[j] -= cum[j] - nptcls
        cnt[j + 1:] = 0
    else:
        cnt[-1] += nptcls - cum[-1]
    ppe[idx] = cnt
    return ppe, int(ppe.sum())


def xgc_initial_coords(mesh, slot_elem, mask, seed=PARTICLE_SEED):
    """test/pseudoXGCm.cpp:224-264 setInitialPtclCoords: two uniforms per SLOT (masked or not) folded
    into the unit triangle, X = A + r1 (B - A) + r2 (C - A) in the slot's row element, z = 0."""
    cap = mask.shape[0]
    r = std_uniform(seed, 2 * cap).reshape(cap, 2)
    x, y = r[:, 0].copy(), r[:, 1].copy()
    f = x + y > 1
    x[f], y[f] = 1 - x[f], 1 - y[f]
    m = mask.astype(bool)
    se = np.where(m, slot_elem, 0)
    V = mesh.coords[mesh.elem2verts[se]]
    pos = (V[:, 0] + x[:, None] * (V[:, 1] - V[:, 0])) + y[:, None] * (V[:, 2] - V[:, 0])
    X = np.zeros((3, cap))
    X[:2, m] = pos[m].T
    return X
