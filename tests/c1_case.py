"""BASELINE configs[0]: test/pseudoPushAndSearch on cube/7k.osh (testing.cmake:106-108:
`pseudoPushAndSearch cube/7k.osh ignored <numPtcls> 156 0 0 1`), restated as a test case.

Seeding (pseudoPushAndSearch.cpp:228-273 setSourceElements): elements up-adjacent to exposed faces
classified on model face 156 share numPtcls evenly, the remainder goes to the highest-numbered
marked element; every particle starts at the centroid of its element (:275-298, Omega_h `average`:
((v0 + v1) + v2 + v3) / 4); push = (maxBBoxLen / 20) * (0, 0, 1) per iteration (:482-496); legacy 3D
search_mesh with maxLoops = 100 (:195-205), updatePtclPositions, rebuild; at most 30 iterations,
stopping when no particle remains (:513-542).
"""
import numpy as np

from meshes import load_fixture

MDL_FACE = 156
NUM_ITERATIONS = 30
MAX_LOOPS = 100


def setup(num_ptcls):
    mesh = load_fixture("cube7k")
    uses = np.bincount(mesh.elem2sides.ravel(), minlength=mesh.nsides)
    exposed = uses == 1
    on_face = exposed & (mesh.side_class_id == MDL_FACE)
    marked = on_face[mesh.elem2sides].any(axis=1)                 # mark_up(mesh, 2, 3, isClassOnFace)
    idx = np.flatnonzero(marked)
    ppe = np.zeros(mesh.nelems, np.int32)
    ppe[idx] = num_ptcls // len(idx)
    ppe[idx[-1]] += num_ptcls % len(idx)
    V = mesh.coords[mesh.elem2verts]                              # [ne, 4, 3]
    centroid = (((V[:, 0] + V[:, 1]) + V[:, 2]) + V[:, 3]) / 4
    ext = (mesh.coords.max(axis=0) - mesh.coords.min(axis=0)).max()
    return mesh, ppe, centroid, ext / 20, (0.0, 0.0, 1.0), idx
