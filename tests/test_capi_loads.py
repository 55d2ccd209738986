"""CPU-only: the C-ABI library loads without a GPU and exports every symbol the header declares."""
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pumipic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    pp = importlib.import_module("pumi-pic_b200")
    L = pp.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
        assert name in pp.capi.PROTOTYPES, "no ctypes prototype for %s" % name
    assert L.pp_build_arch() == b"sm_100a"


def test_product_library_does_not_link_the_oracle():
    import subprocess
    pp = importlib.import_module("pumi-pic_b200")
    out = subprocess.run(["nm", "-D", pp.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in out
    ldd = subprocess.run(["ldd", pp.capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_host_mesh_utilities_match_numpy_twins():
    import meshes
    pp = importlib.import_module("pumi-pic_b200")
    c, ev = pp.host_kuhn_cube(7)
    m = meshes.kuhn_cube(7)
    assert np.array_equal(ev, m.elem2verts) and np.array_equal(c, m.coords)
    e2s, s2v = pp.host_derive_sides(3, ev)
    assert np.array_equal(e2s, m.elem2sides) and np.array_equal(s2v, m.side2verts)
    c2, ev2 = pp.host_plate(9)
    p = meshes.plate(9)
    e2s2, s2v2 = pp.host_derive_sides(2, ev2)
    assert np.array_equal(ev2, p.elem2verts) and np.array_equal(c2, p.coords)
    assert np.array_equal(e2s2, p.elem2sides) and np.array_equal(s2v2, p.side2verts)


def test_compute_call_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    pp = importlib.import_module("pumi-pic_b200")
    import ctypes as C
    m = meshes_small()
    d = pp.capi.MeshDesc(3, m.coords.shape[0], m.elem2verts.shape[0], m.side2verts.shape[0],
                         m.coords.ctypes.data, m.elem2verts.ctypes.data, m.elem2sides.ctypes.data,
                         m.side2verts.ctypes.data, 0, pp.capi.PP_HOST)
    h = C.c_void_p()
    rc = pp.lib().pp_mesh_create(C.byref(d), None, C.byref(h))
    assert rc == 2 and b"failed" in pp.lib().pp_last_error()    # PP_ERR_CUDA, no silent fallback


def meshes_small():
    import meshes
    return meshes.kuhn_cube(2)


def test_mirror_header_compiles_on_the_host_and_input_follows_the_reference(tmp_path):
    """The C++ mirror (pumi-pic_b200/cpp/pumipic_b200.hpp) is usable from a plain g++ translation unit
    (set-up code of an application), and pumipic::Input behaves like pumipic_input.cpp:94-110."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.isdir(cuda_inc):
        pytest.skip("g++ or the CUDA headers are not here")
    importlib.import_module("pumi-pic_b200").lib()          # makes sure the library is built
    exe = str(tmp_path / "input_host")
    libdir = os.path.join(root, "pumi-pic_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(root, "include"), "-I", os.path.join(libdir, "cpp"),
                           "-I", cuda_inc, os.path.join(root, "tests", "cpp", "input_host.cpp"), "-o", exe,
                           "-L", libdir, "-lpumipic_b200", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, timeout=120)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    out = r.stdout.decode()
    assert "pumipic buffer method MINIMUM" in out and "pumipic safe method MINIMUM" in out
