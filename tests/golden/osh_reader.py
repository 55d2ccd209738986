"""Minimal reader for Omega_h `.osh` (format version 9) mesh directories.

Test infrastructure only: it turns the reference's mesh fixtures
(`pumipic-data/*.osh`, read in the authoring container) into small `.npz`
files under tests/golden/ so that the parity tests can run on a box where
/root/reference does not exist.  Layout follows SURVEY.md App. B (the format
itself belongs to Omega_h, which is not vendored in the reference tree).
"""
import struct
import zlib
import numpy as np

_TYPES = {0: np.int8, 2: np.int32, 3: np.int64, 5: np.float64}


class _Stream:
    def __init__(self, buf):
        self.b = buf
        self.p = 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += n
        return v[0]

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def array(self, dtype, compressed):
        n = self.take("i")
        dt = np.dtype(dtype)
        if compressed:
            nbytes = self.take("q")
            data = zlib.decompress(self.raw(nbytes))
        else:
            data = self.raw(n * dt.itemsize)
        a = np.frombuffer(data, dtype=dt.newbyteorder("<"), count=n)
        return a.astype(dt)


def _align(tup, code):
    """Express a down-entity's canonical vertex tuple in its parent's frame."""
    n = len(tup)
    flip = code & 1
    rot = (code >> 1) & 3
    out = [None] * n
    for j in range(n):
        out[(j + rot) % n] = tup[j]
    if flip and n == 3:
        out[1], out[2] = out[2], out[1]
    return out


def read_osh(path):
    """Return a dict with Omega_h-numbered arrays of the serial mesh at `path`."""
    with open(path + "/0.osh", "rb") as f:
        s = _Stream(f.read())
    assert s.take("B") == 0xA1 and s.take("B") == 0x1A, "bad magic"
    compressed = s.take("b")
    family = s.take("b")
    dim = s.take("b")
    s.take("i"); s.take("i")          # comm size / rank
    s.take("b"); s.take("i")          # parting, ghost layers
    have_hints = s.take("b")
    assert family == 0 and have_hints == 0
    nverts = s.take("i")
    down, codes = {}, {}
    for d in range(1, dim + 1):
        down[d] = s.array(np.int32, compressed)
        if d > 1:
            codes[d] = s.array(np.int8, compressed)
    tags = {}
    for d in range(dim + 1):
        ntags = s.take("i")
        for _ in range(ntags):
            nl = s.take("i")
            name = s.raw(nl).decode()
            ncomps = s.take("b")
            typ = s.take("b")
            tags[(d, name)] = (ncomps, s.array(_TYPES[typ], compressed))
    out = {"dim": dim, "nverts": nverts}
    out["coords"] = tags[(0, "coordinates")][1].reshape(nverts, dim).copy()
    edge2verts = down[1].reshape(-1, 2)
    out["edge2verts"] = edge2verts
    face2edges = down[2].reshape(-1, 3)
    fcodes = codes[2].reshape(-1, 3)
    nfaces = face2edges.shape[0]
    # triangle vertices from the aligned edges: e0=(v0,v1), e1=(v1,v2), e2=(v2,v0)
    e0 = edge2verts[face2edges[:, 0]]
    e1 = edge2verts[face2edges[:, 1]]
    r0 = ((fcodes[:, 0] >> 1) & 3) == 1
    r1 = ((fcodes[:, 1] >> 1) & 3) == 1
    v0 = np.where(r0, e0[:, 1], e0[:, 0])
    v1 = np.where(r0, e0[:, 0], e0[:, 1])
    v2 = np.where(r1, e1[:, 0], e1[:, 1])
    v1b = np.where(r1, e1[:, 1], e1[:, 0])
    assert np.array_equal(v1, v1b), "edge alignment inconsistent"
    face2verts = np.stack([v0, v1, v2], axis=1).astype(np.int32)
    out["face2edges"] = face2edges
    out["face2verts"] = face2verts
    if dim == 3:
        tet2faces = down[3].reshape(-1, 4)
        tcodes = codes[3].reshape(-1, 4)
        ntets = tet2faces.shape[0]
        tv = np.empty((ntets, 4), np.int32)
        for t in range(ntets):
            f0 = _align(list(face2verts[tet2faces[t, 0]]), int(tcodes[t, 0]))
            f1 = _align(list(face2verts[tet2faces[t, 1]]), int(tcodes[t, 1]))
            # face 0 = (v0,v2,v1), face 1 = (v0,v1,v3)
            a, c, b = f0
            assert f1[0] == a and f1[1] == b, "tet alignment inconsistent"
            tv[t] = (a, b, c, f1[2])
        out["elem2verts"] = tv
        out["elem2sides"] = tet2faces.astype(np.int32)
        out["side2verts"] = face2verts
    else:
        out["elem2verts"] = face2verts
        out["elem2sides"] = face2edges.astype(np.int32)
        out["side2verts"] = edge2verts.astype(np.int32)
    for d in range(dim + 1):
        if (d, "class_id") in tags:
            out["class_id_%d" % d] = tags[(d, "class_id")][1].astype(np.int32)
        if (d, "class_dim") in tags:
            out["class_dim_%d" % d] = tags[(d, "class_dim")][1].astype(np.int8)
    return out


def read_osh_tags(path):
    """Like read_osh, plus every tag: dict dim, nents[d], down[d], codes[d], verts[d], tags[(d, name)]."""
    with open(path + "/0.osh", "rb") as f:
        s = _Stream(f.read())
    assert s.take("B") == 0xA1 and s.take("B") == 0x1A, "bad magic"
    compressed = s.take("b")
    s.take("b")
    dim = s.take("b")
    s.take("i"); s.take("i"); s.take("b"); s.take("i")
    assert s.take("b") == 0
    nents = [s.take("i"), 0, 0, 0]
    down, codes = {}, {}
    for d in range(1, dim + 1):
        down[d] = s.array(np.int32, compressed)
        nents[d] = len(down[d]) // (d + 1)
        if d > 1:
            codes[d] = s.array(np.int8, compressed)
    tags = {}
    for d in range(dim + 1):
        for _ in range(s.take("i")):
            nl = s.take("i")
            name = s.raw(nl).decode()
            s.take("b")
            typ = s.take("b")
            tags[(d, name)] = s.array(_TYPES[typ], compressed)
    geo = read_osh(path)
    verts = {1: geo["edge2verts"].ravel(), 2: geo["face2verts"].ravel()}
    if dim == 3:
        verts[3] = geo["elem2verts"].ravel()
    return {"dim": dim, "nents": nents, "down": down, "codes": codes, "verts": verts, "tags": tags}


def read_ppm(path):
    """pumipic::write's per-rank record (src/pumipic_file.cpp:82-115), format versions 1 and 2."""
    with open(path, "rb") as f:
        s = _Stream(f.read())
    version = s.take("b")
    out = {"version": version, "full": s.take("b"), "dims": []}
    for _ in range(4):
        o = {"num_entites": s.take("q") if version >= 2 else 0, "num_cores": s.take("i")}
        for k in ("buffered_parts", "offset_ents_per_rank", "ent_to_comm_arr_index",
                  "is_complete_part"):
            o[k] = s.array(np.int32, True)
        o["num_bounds"] = s.take("i")
        o["num_boundaries"] = s.take("i")
        for k in ("boundary_parts", "offset_bounded", "bounded_ent_ids"):
            o[k] = s.array(np.int32, True)
        out["dims"].append(o)
    assert s.p == len(s.b), "trailing bytes in .ppm"
    return out
