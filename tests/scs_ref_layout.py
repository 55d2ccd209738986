"""The Sell-C-sigma geometry the REFERENCE builds for a given particles-per-element array:
oracle/_ref's ref_scs_layout = the reference's chooseChunkHeight / constructChunks /
constructOffsets (particle_structs/src/scs/SCS_buildFns.h:4-153) compiled unmodified."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpumipic_ref_primitives.so")
ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(REF_LIB)


def layout(ppe, max_c=32, sigma=0x7fffffff, V=1024, shuffle_padding=0.1, pad_strat=0):
    L = C.CDLL(REF_LIB)
    ppe = np.ascontiguousarray(ppe, np.int32)
    ne = ppe.shape[0]
    bound = int(ne + (int(ppe.max()) * 3 if ne else 0) // max(V, 1) * ne // max(1, 1) + 16) if ne else 16
    bound = min(max(bound, 4 * ne + 16), 50_000_000)
    cw = np.zeros(ne + 1, np.int32)
    off = np.zeros(bound + 1, np.int32)
    s2c = np.zeros(bound + 1, np.int32)
    c, nch, nsl, nempty = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    cap = L.ref_scs_layout(ne, ppe.ctypes.data_as(ip), max_c, sigma, V, C.c_double(shuffle_padding), pad_strat,
                           C.byref(c), C.byref(nch), cw.ctypes.data_as(ip), C.byref(nsl), off.ctypes.data_as(ip),
                           s2c.ctypes.data_as(ip), bound, C.byref(nempty))
    assert nsl.value <= bound
    return {"capacity": cap, "C": c.value, "nchunks": nch.value, "nslices": nsl.value,
            "chunk_widths": cw[:nch.value].copy(), "offsets": off[:nsl.value + 1].copy(),
            "slice_to_chunk": s2c[:nsl.value].copy(), "num_empty": nempty.value}
