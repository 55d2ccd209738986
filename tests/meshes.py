"""Mesh helpers for the tests: golden fixtures + numpy twins of the host-side generators.

The numpy derivations here are deliberately independent of the C++ host code in
pumi-pic_b200/csrc/pp_host_mesh.cpp so that the two can be checked against each other.
"""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TET_FACE = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
TRI_EDGE = np.array([[0, 1], [1, 2], [2, 0]])


class Mesh:
    """Plain container: dim, coords[nverts,dim], elem2verts, elem2sides, side2verts (+class ids)."""

    def __init__(self, dim, coords, elem2verts, elem2sides, side2verts, class_id=None,
                 side_class_id=None):
        self.dim = int(dim)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.elem2verts = np.ascontiguousarray(elem2verts, dtype=np.int32)
        self.elem2sides = np.ascontiguousarray(elem2sides, dtype=np.int32)
        self.side2verts = np.ascontiguousarray(side2verts, dtype=np.int32)
        self.class_id = None if class_id is None else np.ascontiguousarray(class_id, np.int32)
        self.side_class_id = (None if side_class_id is None
                              else np.ascontiguousarray(side_class_id, np.int32))

    @property
    def nverts(self):
        return self.coords.shape[0]

    @property
    def nelems(self):
        return self.elem2verts.shape[0]

    @property
    def nsides(self):
        return self.side2verts.shape[0]


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, "mesh_%s.npz" % name))
    dim = int(z["dim"])
    return Mesh(dim, z["coords"], z["elem2verts"], z["elem2sides"], z["side2verts"],
                z["class_id_%d" % dim], z["class_id_%d" % (dim - 1)])


def derive_sides(elem2verts, dim):
    """Sides numbered by lexicographic order of their sorted vertex tuple; a side's own vertex
    order is the template order seen from its lowest-numbered adjacent element."""
    tmpl = TET_FACE if dim == 3 else TRI_EDGE
    ev = np.asarray(elem2verts, dtype=np.int64)
    sv = ev[:, tmpl]                       # [ne, nsides_per_elem, dim]
    flat = sv.reshape(-1, dim)
    key = np.sort(flat, axis=1)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    side2verts = flat[first].astype(np.int32)
    elem2sides = inv.reshape(ev.shape[0], dim + 1).astype(np.int32)
    return elem2sides, side2verts


def kuhn_cube(n, length=1.0):
    """n^3 cubes on [0,length]^3, each cut into 6 positively oriented tets (Kuhn / Freudenthal)."""
    g = np.arange(n + 1, dtype=np.float64) * (length / n)
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    coords = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    s = n + 1

    def vid(i, j, k):
        return i + s * (j + s * k)

    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ci, cj, ck = ci.transpose(2, 1, 0).ravel(), cj.transpose(2, 1, 0).ravel(), ck.transpose(2, 1, 0).ravel()
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = np.empty((ci.size, 6, 4), dtype=np.int32)
    for p, perm in enumerate(perms):
        o = np.stack([ci, cj, ck], axis=1)
        pts = [o.copy()]
        for ax in perm:
            o = o.copy()
            o[:, ax] += 1
            pts.append(o)
        v = [vid(q[:, 0], q[:, 1], q[:, 2]) for q in pts]
        sign = np.linalg.det(np.eye(3)[list(perm)])
        if sign < 0:
            v[1], v[2] = v[2], v[1]
        tets[:, p, :] = np.stack(v, axis=1)
    elem2verts = tets.reshape(-1, 4)
    elem2sides, side2verts = derive_sides(elem2verts, 3)
    return Mesh(3, coords, elem2verts, elem2sides, side2verts,
                np.ones(elem2verts.shape[0], np.int32))


def plate(n, length=1.0):
    """n^2 squares on [0,length]^2, each cut into 2 counter-clockwise triangles."""
    g = np.arange(n + 1, dtype=np.float64) * (length / n)
    yy, xx = np.meshgrid(g, g, indexing="ij")
    coords = np.stack([xx.ravel(), yy.ravel()], axis=1)
    s = n + 1
    cj, ci = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ci, cj = ci.ravel(), cj.ravel()
    v00 = ci + s * cj
    v10 = v00 + 1
    v01 = v00 + s
    v11 = v01 + 1
    tris = np.empty((ci.size, 2, 3), dtype=np.int32)
    tris[:, 0, :] = np.stack([v00, v10, v11], axis=1)
    tris[:, 1, :] = np.stack([v00, v11, v01], axis=1)
    elem2verts = tris.reshape(-1, 3)
    elem2sides, side2verts = derive_sides(elem2verts, 2)
    return Mesh(2, coords, elem2verts, elem2sides, side2verts,
                np.ones(elem2verts.shape[0], np.int32))
