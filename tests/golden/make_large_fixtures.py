"""Large mesh fixtures for the benchmark tools (NOT committed: tests/golden/_large/ is git-ignored, but
it travels to the GPU box with the snapshot like the built libraries).  Run in the authoring container
(needs /root/reference); __graft_entry__.build() calls it when the fixture is missing.

    xgc/2M.osh -> tests/golden/_large/mesh_xgc2M.npz   (BASELINE configs[3]: 2 009 137 triangles)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
SRC = "/root/reference/pumipic-data/xgc/2M.osh"
OUT = os.path.join(HERE, "_large", "mesh_xgc2M.npz")
KEEP = ("coords", "elem2verts", "elem2sides", "side2verts", "class_id_2", "class_id_1")


def make(force=False):
    if os.path.exists(OUT) and not force:
        return OUT
    if not os.path.exists(SRC):
        return None
    from osh_reader import read_osh
    m = read_osh(SRC)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, dim=np.int32(m["dim"]), **{k: m[k] for k in KEEP})
    return OUT


if __name__ == "__main__":
    p = make(force="-f" in sys.argv)
    print(p, os.path.getsize(p) if p else None)
