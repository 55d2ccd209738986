"""Shared helpers of the GPU parity tests (they all call through the C ABI)."""
import importlib

import numpy as np


def pp():
    return importlib.import_module("pumi-pic_b200")


def torch():
    import torch as t
    return t


def dev(a):
    t = torch()
    return t.as_tensor(np.ascontiguousarray(a)).cuda()


def make_gpu_mesh(mesh):
    return pp().Mesh(mesh.dim, mesh.coords, mesh.elem2verts, mesh.elem2sides, mesh.side2verts,
                     mesh.class_id)


PARTICLE = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]  # test_adj.cpp:23


def make_ps(kind, ppe, **kw):
    return pp().ParticleStructure(kind, PARTICLE, ppe, **kw)
