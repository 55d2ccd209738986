"""Canonical, order-insensitive-where-the-reference-is form of a PICpart, shared by the fixture
generator (reference files decoded in Python) and the parity test (product library output).

`mesh`: dict with dim, nents[d], down[d] (d -> d-1, rows), codes[2], tags[(d, name)]
`ppm` : dict version, full, dims[d] = {num_entites, num_cores, buffered_parts, ...}

What is canonicalised and why:
  * edges are digested with sorted rows + a packed orientation bitmask: the fixtures' edges on
    the outer boundary of a PICpart are stored reversed (they were derived from the faces that
    are present), the current reference copies them (part_construct.cpp:546-552); the test
    checks that only such edges differ.
  * rank-local ids of entities of a part of which only a boundary is held come from atomics
    (pumipic_comm.cpp:66-75): comm-array indices of those entities and the matching
    bounded_ent_ids lists are compared as sorted sets.
  * sbar ids are numbered in the iteration order of a std::unordered_map (pumipic_lb.cpp:200-
    212, implementation-defined): the element -> sbar map is compared up to relabelling (ids
    replaced by order of first appearance) together with the parts of every sbar.
"""
import hashlib

import numpy as np


def sha(a, dtype):
    return hashlib.sha1(np.ascontiguousarray(a, dtype).tobytes()).hexdigest()


def relabel(ids):
    """ids replaced by their order of first appearance."""
    ids = np.asarray(ids)
    _, first, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(np.argsort(first))
    return order[inv]


def canon_picpart(mesh, ppm, nranks):
    dim = mesh["dim"]
    out = {"dim": dim, "is_full_mesh": int(ppm["full"]), "dims": []}
    for d in range(dim + 1):
        n = int(mesh["nents"][d])
        t = lambda name: mesh["tags"][(d, name)]   # noqa: E731
        own = np.asarray(t("ownership"))
        e = {"nents": n,
             "ownership": sha(own, np.int32), "gids": sha(t("gids"), np.int64),
             "rank_lids": sha(t("rank_lids"), np.int32), "class_id": sha(t("class_id"), np.int32),
             "class_dim": sha(t("class_dim"), np.int8)}
        if d == 0:
            e["coordinates"] = sha(t("coordinates"), np.float64)
        if d == 1:
            rows = np.asarray(mesh["down"][1]).reshape(-1, 2)
            e["edge_verts_sorted"] = sha(np.sort(rows, axis=1), np.int32)
            e["edge_orientation_bits"] = np.packbits(rows[:, 0] > rows[:, 1]).tobytes().hex()
        if d == 2:
            e["down"] = sha(mesh["down"][2], np.int32)
            e["ent2verts"] = sha(mesh["verts"][2], np.int32)
        if d == dim:
            e["safe"] = sha(t("safe"), np.int32)
            sb = np.asarray(t("sbar_id"))
            e["sbar_partition"] = sha(relabel(sb), np.int32)
        p = ppm["dims"][d]
        for k in ("num_entites", "num_cores", "num_bounds", "num_boundaries"):
            e[k] = int(p[k])
        for k in ("buffered_parts", "offset_ents_per_rank", "is_complete_part", "boundary_parts",
                  "offset_bounded"):
            e[k] = [int(x) for x in p[k]]
        comp = np.asarray(p["is_complete_part"])
        cai = np.asarray(p["ent_to_comm_arr_index"])
        bnd = comp[own] == 1
        e["comm_index_complete"] = sha(cai[~bnd], np.int32)
        e["comm_index_boundary_sorted"] = sha(np.sort(cai[bnd]), np.int32)
        off = p["offset_bounded"]
        ids = np.asarray(p["bounded_ent_ids"])
        segs = [np.sort(ids[off[s]:off[s + 1]]) for s in range(len(off) - 1)]
        e["bounded_ent_ids_sorted"] = sha(np.concatenate(segs) if segs else ids, np.int32)
        out["dims"].append(e)
    return out
