"""Regenerate the PICpart golden fixtures from the reference's own output files.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_picpart_fixtures.py

Source: pumipic-data/xgc/{24k,120k}_4.ppm -- what the reference's `file_rw` test wrote
(test/test_file.cpp, test/testing.cmake:60-78) for the 4-part class partitions
xgc/{24k,120k}_4.cpn.  Everything is decoded with the independent Python reader in this
directory (osh_reader.py), never with the product library.

Outputs (committed):
  mesh_xgc120k.npz            input mesh (coords, elements, classification) -- the 24k one exists
  picpart_xgc{24k,120k}_4.json expected PICparts: small arrays in full, large arrays as sha1
                               digests of a canonical form (picpart_canon.py, shared with the test)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from osh_reader import read_osh, read_osh_tags, read_ppm   # noqa: E402
from picpart_canon import canon_picpart                    # noqa: E402

DATA = "/root/reference/pumipic-data/xgc"
# buffer / safe methods that reproduce each fixture (0 FULL, 1 BFS): 24k_4 holds full-mesh
# PICparts with a BFS safe zone, 120k_4 is `bfs bfs` as in testing.cmake:73-78
CASES = {"24k": (0, 1), "120k": (1, 1)}


def read_cpn(path):
    toks = open(path).read().split()
    size = int(toks[0])
    owners = np.zeros(size + 1, np.int32)
    for c, o in zip(toks[1::2], toks[2::2]):
        owners[int(c)] = int(o)
    return owners


if __name__ == "__main__":
    m = read_osh(os.path.join(DATA, "120k.osh"))
    keep = {k: m[k] for k in ("coords", "elem2verts", "class_id_0", "class_dim_0", "class_id_1",
                              "class_dim_1", "class_id_2", "class_dim_2")}
    out = os.path.join(HERE, "mesh_xgc120k.npz")
    np.savez_compressed(out, **keep)
    print("mesh_xgc120k.npz", os.path.getsize(out), "bytes")
    for name, (bm, sm) in CASES.items():
        nranks = 4
        exp = {"source": "pumipic-data/xgc/%s_%d.ppm" % (name, nranks), "nranks": nranks,
               "buffer_method": bm, "safe_method": sm,
               "class_owner": read_cpn(os.path.join(DATA, "%s_%d.cpn" % (name, nranks))).tolist(),
               "ranks": []}
        meshes = []
        for r in range(nranks):
            base = os.path.join(DATA, "%s_%d.ppm" % (name, nranks), "%s_%d" % (name, r))
            mesh = read_osh_tags(base + ".osh")
            ppm = read_ppm(base + ".ppm")
            meshes.append(mesh)
            exp["ranks"].append(canon_picpart(mesh, ppm, nranks))
        # the parts of every sbar: an element's sbar = its owner + every part it is safe on
        dim = meshes[0]["dim"]
        safe_on = [set(m["tags"][(dim, "gids")][m["tags"][(dim, "safe")] != 0].tolist()) for m in meshes]
        for r, m in enumerate(meshes):
            gid, own, sb = (m["tags"][(dim, k)] for k in ("gids", "ownership", "sbar_id"))
            table = {}
            for i in np.unique(sb, return_index=True)[1]:
                parts = sorted({int(own[i])} | {b for b in range(nranks) if int(gid[i]) in safe_on[b]})
                if r in parts:
                    table[int(sb[i])] = parts
            exp["ranks"][r]["sbars"] = {str(k): v for k, v in sorted(table.items())}
        out = os.path.join(HERE, "picpart_xgc%s_%d.json" % (name, nranks))
        with open(out, "w") as f:
            json.dump(exp, f, indent=1, sort_keys=True)
        print(os.path.basename(out), os.path.getsize(out), "bytes")
