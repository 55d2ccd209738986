"""The C++ mirror of the PUMI-PIC API (pumi-pic_b200/cpp/pumipic_b200.hpp) driven like the
reference's test/pseudoPushAndSearch.cpp: user push lambda through ps::parallel_for, search_mesh,
updatePtclPositions lambda, migrate_lb_ptcls -- compared with the CPU oracle's serial loop by
particle id (element ids and positions bit-exact)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_api as orc
import ptcl_init as pi
from meshes import kuhn_cube

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin", "pseudo_push_and_search")


def _build_driver():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pp_build", os.path.join(ROOT, "pumi-pic_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_cpp_tests()


def test_cpp_mirror_header_is_plain_cxx():
    """host-only translation units (g++, no nvcc) can include the mirror and the C ABI header"""
    src = '#include "pumipic_b200.hpp"\nint main(){ pumipic::TeamPolicy p(1,32); return p.team_size()==32?0:1; }\n'
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "pumi-pic_b200", "cpp"), "-I", "/usr/local/cuda/include", "-"],
                       input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [0, 1, 3])
def test_cpp_driver_matches_oracle(tmp_path, kind):
    if not os.path.exists(BIN):
        _build_driver()
    N, nptcls, nsteps = 6, 30000, 6
    mesh = kuhn_cube(N)
    om = orc.OracleMesh(mesh)
    ppe = pi.even_ppe(mesh.nelems, nptcls)
    pel = np.repeat(np.arange(mesh.nelems, dtype=np.int32), ppe)
    X, D = pi.init3d_internal(mesh, pel, np.ones(nptcls, np.uint8))
    dist = 1.3 * pi.push_distance(mesh)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("ii", N, nptcls)); f.write(struct.pack("d", dist))
        f.write(ppe.astype(np.int32).tobytes()); f.write(pel.tobytes())
        f.write(np.ascontiguousarray(X).tobytes()); f.write(np.ascontiguousarray(D).tobytes())
    r = subprocess.run([BIN, fin, fout, str(nsteps), str(kind)], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stdout.decode() + r.stderr.decode()
    raw = open(fout, "rb").read()
    n = struct.unpack("i", raw[:4])[0]
    rec = np.frombuffer(raw[4:], dtype=np.dtype([("pid", "<i4"), ("elem", "<i4"), ("pos", "<f8", 3)]), count=n)
    # oracle loop
    pid = np.arange(nptcls); se = pel.copy(); Xo = X.copy(); Do = D.copy()
    for _ in range(nsteps):
        if len(pid) == 0:
            break
        To = Xo + dist * Do
        found, ids, _, _, st = om.search_mesh(se, np.ones(len(pid), np.uint8), Xo, To, looplimit=100)
        assert found
        keep = ids >= 0
        pid, se, Xo, Do = pid[keep], ids[keep].astype(np.int32), To[:, keep].copy(), Do[:, keep].copy()
    assert n == len(pid)
    order = np.argsort(rec["pid"])
    assert np.array_equal(rec["pid"][order], pid)
    assert np.array_equal(rec["elem"][order], se)
    assert np.array_equal(rec["pos"][order], Xo.T)      # bit-exact positions
