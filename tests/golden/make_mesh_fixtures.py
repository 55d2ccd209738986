"""Regenerate tests/golden/mesh_*.npz from the reference's `.osh` fixtures.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_mesh_fixtures.py
The `.npz` files are committed; the GPU box never reads /root/reference.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from osh_reader import read_osh  # noqa: E402

DATA = "/root/reference/pumipic-data"
FIXTURES = {
    "cube7k": "cube/7k.osh",
    "tri8_parDiag": "plate/tri8_parDiag.osh",
    "tri8": "plate/tri8.osh",
    "xgc24k": "xgc/24k.osh",
}

def msh_tets(path):
    """Nodes + tets (gmsh v2.2 element type 4, file order) of a .msh file: geometry only."""
    lines = open(path).read().split("\n")
    i = lines.index("$Nodes")
    n = int(lines[i + 1])
    coords = np.array([[float(t) for t in lines[i + 2 + k].split()[1:4]] for k in range(n)])
    j = lines.index("$Elements")
    tets = []
    for k in range(int(lines[j + 1])):
        t = lines[j + 2 + k].split()
        if int(t[1]) == 4:
            nt = int(t[2])
            tets.append([int(v) - 1 for v in t[3 + nt:3 + nt + 4]])
    return coords, np.array(tets, np.int32)


if __name__ == "__main__":
    c, t = msh_tets(os.path.join(DATA, "cube6tet.msh"))
    np.savez_compressed(os.path.join(HERE, "cube6tet.npz"), coords=c, tets=t)
    for name, rel in FIXTURES.items():
        m = read_osh(os.path.join(DATA, rel))
        keep = {k: v for k, v in m.items() if isinstance(v, np.ndarray)}
        keep["dim"] = np.int32(m["dim"])
        out = os.path.join(HERE, "mesh_%s.npz" % name)
        np.savez_compressed(out, **keep)
        print(name, m["dim"], m["elem2verts"].shape, os.path.getsize(out), "bytes")
