"""CPU-only parity of the set-up side against the reference's own
files: PICpart construction (part_construct.cpp:73-262), setupComm (pumipic_comm.cpp:12-184), the
sbar regions of ParticleBalancer (pumipic_lb.cpp:23-82), Omega_h's entity derivation and `.osh` /
`.ppm` I/O (pumipic_file.cpp:45-205).

Golden data: pumipic-data/xgc/{24k,120k}.osh and the 4-rank PICparts the reference wrote from them
(xgc/{24k,120k}_4.ppm), decoded by the independent Python reader in tests/golden/ into
picpart_xgc*_4.json (tests/golden/make_picpart_fixtures.py).  Integer work: bit-exact, except
where the reference itself is order-free (see tests/golden/picpart_canon.py).
"""
import importlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
from picpart_canon import canon_picpart  # noqa: E402

pp = importlib.import_module("pumi-pic_b200")

MESH_NPZ = {"24k": "mesh_xgc24k.npz", "120k": "mesh_xgc120k.npz"}


def _full_mesh(name):
    """The fixture's full mesh rebuilt from coordinates + elements: entities derived our way."""
    d = np.load(os.path.join(GOLD, MESH_NPZ[name]))
    m = pp.HostMesh.from_elems(2, d["coords"], d["elem2verts"])
    for k in range(3):
        m.set_tag(k, "class_dim", d["class_dim_%d" % k].astype(np.int8))
        m.set_tag(k, "class_id", d["class_id_%d" % k].astype(np.int32))
    return m, d


def _as_canon_input(picpart):
    """A product PICpart in the dict form tests/golden/picpart_canon.py digests."""
    m = picpart.mesh()
    dim = m.dim
    mesh = {"dim": dim, "nents": [m.nents(d) for d in range(4)],
            "down": {d: m.down(d).ravel() for d in range(1, dim + 1)},
            "verts": {d: m.ent2verts(d).ravel() for d in range(1, dim + 1)},
            "tags": {(d, n): m.tag(d, n).ravel() for d in range(dim + 1) for n in m.tag_names(d)}}
    ppm = {"full": int(picpart.is_full_mesh), "dims": []}
    for d in range(4):
        i = picpart.dim_info(d)
        i["num_entites"] = i["num_entities"]
        ppm["dims"].append(i)
    return mesh, ppm


def _picpart_boundary_edges(m):
    """Edges with exactly one adjacent face in this mesh."""
    uses = np.bincount(m.down(2).ravel(), minlength=m.nents(1))
    return uses == 1


@pytest.mark.parametrize("name", ["24k", "120k"])
def test_derived_entities_match_omega_h(name):
    """Edges derived from the triangles are the ones Omega_h stored in xgc/*.osh: numbering,
    vertex order and alignment codes (the committed fixture keeps Omega_h's arrays)."""
    m, d = _full_mesh(name)
    if "edge2verts" in d.files:
        assert np.array_equal(m.ent2verts(1), d["edge2verts"])
        assert np.array_equal(m.down(2), d["face2edges"])
    assert np.array_equal(m.ent2verts(2), d["elem2verts"])
    # classification arrays only line up if the numbering is Omega_h's
    assert m.nents(1) == d["class_id_1"].shape[0]
    ref = "/root/reference/pumipic-data/xgc/%s.osh" % name
    if os.path.isdir(ref):
        f = pp.HostMesh.read_osh(ref)
        for k in (1, 2):
            assert np.array_equal(f.down(k), m.down(k))
            assert np.array_equal(f.ent2verts(k), m.ent2verts(k))
        assert np.array_equal(f.codes(2), m.codes(2))
        assert np.array_equal(f.coords(), m.coords())


@pytest.mark.parametrize("name", ["24k", "120k"])
def test_picparts_match_reference_files(name):
    exp = json.load(open(os.path.join(GOLD, "picpart_xgc%s_4.json" % name)))
    full, d = _full_mesh(name)
    class_owner = np.asarray(exp["class_owner"], np.int32)
    owner = class_owner[d["class_id_2"]]          # setOwnerByClassification
    nranks = exp["nranks"]
    mine_all, want_all, max_sbars = {}, {}, set()
    for r in range(nranks):
        mine = pp.Picpart.build(full, owner, nranks, r, exp["buffer_method"], exp["safe_method"])
        got = canon_picpart(*_as_canon_input(mine), nranks)
        want = exp["ranks"][r]
        assert got["is_full_mesh"] == want["is_full_mesh"]
        for k in range(3):
            g, w = got["dims"][k], want["dims"][k]
            for key in w:
                if key == "edge_orientation_bits":
                    continue
                assert g[key] == w[key], "rank %d dim %d: %s differs" % (r, k, key)
        # edges: same numbering and vertices; orientation may differ only on the outer boundary of
        # the PICpart (the fixtures derived those edges from the faces present, the current
        # reference copies the full mesh's orientation, part_construct.cpp:546-552)
        gb = np.unpackbits(np.frombuffer(bytes.fromhex(got["dims"][1]["edge_orientation_bits"]), np.uint8))
        wb = np.unpackbits(np.frombuffer(bytes.fromhex(want["dims"][1]["edge_orientation_bits"]), np.uint8))
        nedges = got["dims"][1]["nents"]
        diff = (gb != wb)[:nedges]
        assert not diff[~_picpart_boundary_edges(mine.mesh())].any()
        if want["is_full_mesh"]:
            assert not diff.any()
        # sbars: same regions (parts), ids spaced by the number of parts and numbered by the
        # smallest part in rank order; the order inside a rank is implementation-defined
        table, max_sbar = mine.sbars()
        assert sorted(table.values()) == sorted(tuple(v) for v in want["sbars"].values())
        for sid, parts in table.items():
            assert r in parts
        mine_all.update(table)
        want_all.update({int(k): tuple(v) for k, v in want["sbars"].items()})
        max_sbars.add(max_sbar)

    # ids numbered by rank q (the smallest part) fill the same contiguous range as in the fixture
    def rng(tab):
        out = {}
        for sid, parts in tab.items():
            lo, hi = out.get(parts[0], (1 << 30, -1))
            out[parts[0]] = (min(lo, sid), max(hi, sid + len(parts)))
        return out
    assert rng(mine_all) == rng(want_all)
    assert max_sbars == {max(hi for _, hi in rng(want_all).values())}


def test_cross_rank_consistency_of_boundary_lists():
    """What rank s holds of rank r's entities without holding all of them must be exactly what r
    lists in bounded_ent_ids for s, in s's comm-array order (pumipic_comm.cpp:107-180)."""
    full, d = _full_mesh("120k")
    exp = json.load(open(os.path.join(GOLD, "picpart_xgc120k_4.json")))
    owner = np.asarray(exp["class_owner"], np.int32)[d["class_id_2"]]
    parts = [pp.Picpart.build(full, owner, 4, r, pp.BFS, pp.BFS) for r in range(4)]
    checked = 0
    for k in (0, 1):
        infos = [p.dim_info(k) for p in parts]
        for s in range(4):
            m = parts[s].mesh()
            own, rl = m.tag(k, "ownership"), m.tag(k, "rank_lids")
            for r in range(4):
                if infos[s]["is_complete_part"][r] != 1:
                    continue
                sel = own == r
                lid = infos[s]["ent_to_comm_arr_index"][sel] - infos[s]["offset_ents_per_rank"][r]
                off = infos[r]["offset_bounded"]
                lst = infos[r]["bounded_ent_ids"][off[s]:off[s + 1]]
                assert len(lst) == sel.sum() > 0
                assert np.array_equal(lst[lid], rl[sel])
                assert s in infos[r]["boundary_parts"]
                checked += 1
    assert checked >= 4


def test_ppm_write_read_round_trip(tmp_path):
    """pumipic::write then pumipic::read (test/test_file.cpp:39-121): every member survives."""
    full, d = _full_mesh("24k")
    exp = json.load(open(os.path.join(GOLD, "picpart_xgc24k_4.json")))
    owner = np.asarray(exp["class_owner"], np.int32)[d["class_id_2"]]
    prefix = str(tmp_path / "xgc24k")
    built = [pp.Picpart.build(full, owner, 4, r, pp.BFS, pp.BFS) for r in range(4)]
    for p in built:
        p.write(prefix)
    for r, p in enumerate(built):
        q = pp.Picpart.read(prefix, 4, r)
        assert q.is_full_mesh == p.is_full_mesh and q.nranks == 4 and q.rank == r
        pm, qm = p.mesh(), q.mesh()
        assert pm.dim == qm.dim
        for k in range(3):
            assert pm.nents(k) == qm.nents(k)
            if k:
                assert np.array_equal(pm.down(k), qm.down(k))
                assert np.array_equal(pm.ent2verts(k), qm.ent2verts(k))
            assert pm.tag_names(k) == qm.tag_names(k)
            for t in pm.tag_names(k):
                assert np.array_equal(pm.tag(k, t), qm.tag(k, t)), t
            a, b = p.dim_info(k), q.dim_info(k)
            for key in a:
                if key == "ent_l2g":
                    assert np.array_equal(a[key], qm.tag(k, "global_serial"))
                else:
                    assert np.array_equal(a[key], b[key]), key
        assert p.sbars()[0] == q.sbars()[0]
    with pytest.raises(pp.PumipicError, match="does not exist"):
        pp.Picpart.read(str(tmp_path / "nothing"), 4, 0)


def test_ppm_uncompressed_files_and_corrupt_input(tmp_path):
    """A reference built without zlib writes raw arrays (src/pumipic_file.cpp:76-80) and the file
    carries no flag: both forms are written on request and read whatever the setting; a file for
    another rank count or with damaged arrays is an error, never an allocation by a bogus count."""
    full, d = _full_mesh("24k")
    exp = json.load(open(os.path.join(GOLD, "picpart_xgc24k_4.json")))
    owner = np.asarray(exp["class_owner"], np.int32)[d["class_id_2"]]
    built = [pp.Picpart.build(full, owner, 4, r, pp.BFS, pp.BFS) for r in range(2)]
    lib = pp.capi.lib()
    raw_prefix, z_prefix = str(tmp_path / "raw"), str(tmp_path / "z")
    try:
        lib.pp_host_ppm_set_compression(0)
        for p in built:
            p.write(raw_prefix)
        lib.pp_host_ppm_set_compression(1)
        for p in built:
            p.write(z_prefix)
        raw_file = os.path.join(raw_prefix + "_4.ppm", "raw_0.ppm")
        z_file = os.path.join(z_prefix + "_4.ppm", "z_0.ppm")
        assert os.path.getsize(raw_file) > os.path.getsize(z_file)
        for setting in (1, 0):                       # either file under either setting
            lib.pp_host_ppm_set_compression(setting)
            for prefix in (raw_prefix, z_prefix):
                for r, p in enumerate(built):
                    q = pp.Picpart.read(prefix, 4, r)
                    for k in range(3):
                        a, b = p.dim_info(k), q.dim_info(k)
                        for key in a:
                            if key != "ent_l2g":
                                assert np.array_equal(a[key], b[key]), key
    finally:
        lib.pp_host_ppm_set_compression(1)
    # the files of a 4-rank run are not a 2-rank PICpart
    os.rename(z_prefix + "_4.ppm", z_prefix + "_2.ppm")
    with pytest.raises(pp.PumipicError, match="not written for 2 ranks"):
        pp.Picpart.read(z_prefix, 2, 0)
    os.rename(z_prefix + "_2.ppm", z_prefix + "_4.ppm")
    # damaged arrays: a huge entry count in front of a short stream, and a truncated file
    blob = bytearray(open(raw_file, "rb").read())
    bad = bytearray(blob)
    bad[14:18] = (0x7fffffff).to_bytes(4, "little")   # first array's count (after version, flag, i64, i32)
    open(raw_file, "wb").write(bytes(bad))
    with pytest.raises(pp.PumipicError, match="truncated or corrupt"):
        pp.Picpart.read(raw_prefix, 4, 0)
    open(raw_file, "wb").write(bytes(blob[: len(blob) // 2]))
    with pytest.raises(pp.PumipicError, match="truncated or corrupt"):
        pp.Picpart.read(raw_prefix, 4, 0)
    open(raw_file, "wb").write(bytes([9]) + bytes(blob[1:]))
    with pytest.raises(pp.PumipicError, match="unsupported version 9"):
        pp.Picpart.read(raw_prefix, 4, 0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/pumipic-data/xgc/120k_4.ppm"),
                    reason="reference data only exists in the authoring container")
def test_reads_reference_ppm_files_directly():
    """Our `.ppm` / `.osh` readers on the reference's files agree with the Python decoder."""
    from osh_reader import read_osh_tags, read_ppm
    base = "/root/reference/pumipic-data/xgc/120k"
    for r in range(4):
        q = pp.Picpart.read(base, 4, r)
        mesh = read_osh_tags("%s_4.ppm/120k_%d.osh" % (base, r))
        ppm = read_ppm("%s_4.ppm/120k_%d.ppm" % (base, r))
        qm = q.mesh()
        for k in range(3):
            assert qm.nents(k) == mesh["nents"][k]
            if k:
                assert np.array_equal(qm.down(k).ravel(), mesh["down"][k])
            for (dd, n), v in mesh["tags"].items():
                if dd == k:
                    assert np.array_equal(qm.tag(k, n).ravel(), v)
            i = q.dim_info(k)
            for key in ("buffered_parts", "offset_ents_per_rank", "ent_to_comm_arr_index",
                        "is_complete_part", "boundary_parts", "offset_bounded", "bounded_ent_ids"):
                assert np.array_equal(i[key], ppm["dims"][k][key]), key
            assert i["num_entities"] == ppm["dims"][k]["num_entites"]


def test_partition_files(tmp_path):
    """`.ptn` and `.cpn` readers (pumipic_input.cpp:44-89) and their error paths."""
    ptn = tmp_path / "four.ptn"
    ptn.write_text("0\n1\n1\n0\n")
    assert pp.host_read_partition(ptn, 4).tolist() == [0, 1, 1, 0]
    with pytest.raises(pp.PumipicError, match="holds 4 owners"):
        pp.host_read_partition(ptn, 5)
    cpn = tmp_path / "c.cpn"
    cpn.write_text("3\n1 0\n2 1\n3 1\n")
    assert pp.host_read_partition(cpn, 4, [1, 3, 2, 1]).tolist() == [0, 1, 1, 0]
    with pytest.raises(pp.PumipicError, match="outside"):
        pp.host_read_partition(cpn, 2, [1, 7])
    with pytest.raises(pp.PumipicError, match="no extension"):
        pp.host_read_partition(tmp_path / "noext", 1)
    bad = tmp_path / "x.txt"
    bad.write_text("0")
    with pytest.raises(pp.PumipicError, match="Only .ptn and .cpn"):
        pp.host_read_partition(bad, 1)
    with pytest.raises(pp.PumipicError, match="Cannot open"):
        pp.host_read_partition(tmp_path / "missing.ptn", 1)


def test_osh_round_trip_3d(tmp_path):
    """A tet mesh derived from elements survives write -> read; every tet's faces and edges carry
    consistent alignment codes (derive_verts re-checks them on read)."""
    coords, elems = pp.host_kuhn_cube(3)
    m = pp.HostMesh.from_elems(3, coords, elems)
    assert m.nents(0) == 64 and m.nents(3) == 162
    # Euler characteristic of a ball: V - E + F - T = 1
    assert m.nents(0) - m.nents(1) + m.nents(2) - m.nents(3) == 1
    assert np.array_equal(m.ent2verts(3), elems)
    m.set_tag(3, "owner", (np.arange(m.nents(3)) % 3).astype(np.int32))
    m.write_osh(tmp_path / "cube.osh")
    q = pp.HostMesh.read_osh(tmp_path / "cube.osh")
    for k in (1, 2, 3):
        assert np.array_equal(m.down(k), q.down(k))
        assert np.array_equal(m.ent2verts(k), q.ent2verts(k))
    assert np.array_equal(m.coords(), q.coords())
    assert np.array_equal(q.tag(3, "owner"), m.tag(3, "owner"))
    # every face is used by one or two tets, and its vertices are the tet's template face
    uses = np.bincount(m.down(3).ravel(), minlength=m.nents(2))
    assert set(np.unique(uses)) <= {1, 2}
    tf = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [2, 0, 3]])
    fv = m.ent2verts(2)
    for t in (0, 57, 161):
        for k in range(4):
            assert sorted(fv[m.down(3)[t, k]]) == sorted(elems[t][tf[k]])


def test_picpart_3d_partial_buffer_properties():
    """3D has no reference file to pin against: structural invariants of a partially buffered
    Kuhn-cube PICpart (test_input_construct.cpp / test_comm_array.cpp style)."""
    coords, elems = pp.host_kuhn_cube(6)
    full = pp.HostMesh.from_elems(3, coords, elems)
    cx = coords[elems].mean(axis=1)[:, 0]
    owner = np.minimum((cx * 4).astype(np.int32), 3)          # 4 slabs along x
    parts = [pp.Picpart.build(full, owner, 4, r, pp.BFS, pp.BFS, 1, 1) for r in range(4)]
    assert not parts[0].is_full_mesh
    tot = np.zeros(4, np.int64)
    for r, p in enumerate(parts):
        m = p.mesh()
        i3 = p.dim_info(3)
        own = m.tag(3, "ownership")
        l2g = i3["ent_l2g"]
        assert np.array_equal(own, owner[l2g]) and np.all(np.diff(l2g) > 0)
        assert np.array_equal(m.tag(3, "global_serial"), l2g)
        # slab r buffers its neighbours only
        assert i3["buffered_parts"].tolist() == [q for q in (r - 1, r + 1) if 0 <= q < 4]
        safe = m.tag(3, "safe")
        assert np.all(safe[own == r] == 1) and safe.sum() > (own == r).sum()
        # geometry travels with the numbering
        assert np.array_equal(m.coords()[m.ent2verts(3)], coords[elems[l2g]])
        # comm-array indices are a permutation per dimension; gids are owner-major
        for k in range(4):
            ik = p.dim_info(k)
            assert sorted(ik["ent_to_comm_arr_index"].tolist()) == list(range(ik["nents"]))
            g = m.tag(k, "gids")
            o = m.tag(k, "ownership")
            assert np.array_equal(g - m.tag(k, "rank_lids"), np.asarray(
                [0] + np.cumsum(np.bincount(full_owner(full, owner, k), minlength=4)).tolist())[o])
        tot[r] = (own == r).sum()
    assert tot.sum() == full.nents(3)


def full_owner(full, owner, k):
    """defineOwners on the full mesh: minimum owner over the adjacent elements."""
    dim = full.dim
    own = owner
    for d in range(dim, k, -1):
        dn = full.down(d)
        nxt = np.full(full.nents(d - 1), 1 << 30, np.int64)
        np.minimum.at(nxt, dn.ravel(), np.repeat(own, d + 1))
        own = nxt
    return np.asarray(own, np.int64)
