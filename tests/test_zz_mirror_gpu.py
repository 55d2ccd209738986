"""GPU tests added when the round's GPU time was spent (hence the file name: last under `-x`):
the self-checking C++ driver tests/cpp/mirror_api.cu (getMemberView, getPIDs, the PICpart-record
Mesh and its accessors, setUnsafeProcs, ParticleBalancer, PS_Comm_* on one rank) and getPIDs
through the C ABI on every structure kind (particle_structs/test/test_structure.cpp:354-378)."""
import os
import subprocess

import numpy as np
import pytest

from gpu_common import pp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin", "mirror_api")


def test_mirror_api_driver():
    if not os.path.exists(BIN):
        import importlib.util
        spec = importlib.util.spec_from_file_location("pp_build", os.path.join(ROOT, "pumi-pic_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_cpp_tests()
    r = subprocess.run([BIN], capture_output=True, timeout=300)
    assert r.returncode == 0 and b"MIRROR_API_OK" in r.stdout, r.stdout.decode() + r.stderr.decode()


@pytest.mark.parametrize("kind", ["scs", "csr", "cabm", "dps"])
@pytest.mark.parametrize("ne,np_", [(50, 1000), (2500, 100000), (7, 0)])
def test_get_pids(kind, ne, np_):
    P = pp()
    K = {"scs": P.capi.PP_PS_SCS, "csr": P.capi.PP_PS_CSR, "cabm": P.capi.PP_PS_CABM,
         "dps": P.capi.PP_PS_DPS}[kind]
    rng = np.random.default_rng(11)
    ppe = np.zeros(ne, np.int32)
    if np_:
        w = rng.random(ne) ** 3
        w[rng.random(ne) < 0.2] = 0
        w[0] += 1e-3
        ppe = np.floor(w / w.sum() * np_).astype(np.int32)
        ppe[0] += np_ - ppe.sum()
    ps = P.ParticleStructure(K, [(np.int32, 1)], ppe)
    pids, offsets = ps.get_pids()
    pids, offsets = pids.cpu().numpy(), offsets.cpu().numpy()
    slot_elem, mask = ps.slot_elem_and_mask()
    assert pids.shape[0] == np_ and offsets.shape[0] == ne + 1
    assert np.array_equal(offsets, np.concatenate([[0], np.cumsum(ppe)]))
    if np_ == 0:
        return
    assert mask[pids].all() and len(np.unique(pids)) == np_
    elem_of_pid = slot_elem[pids]
    assert np.array_equal(elem_of_pid, np.repeat(np.arange(ne), ppe))       # grouped by element
    same = elem_of_pid[1:] == elem_of_pid[:-1]
    assert np.all(np.diff(pids)[same] > 0)                                  # ascending slot inside a group
