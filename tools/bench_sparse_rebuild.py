#!/usr/bin/env python
"""Rebuild of a mostly empty Sell-C-sigma structure on ONE GPU: the structure a rank of an 8-GPU run holds
when its PICpart buffers the whole mesh (8 M rows, particles in its own 1 M).  Times the rebuild with the
split-rows layout (only the non-empty rows are sorted) and without; prints one JSON line per setting with
the library's phase timers."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8_000_000)
    ap.add_argument("--occupied", type=int, default=1_000_000)
    ap.add_argument("--ppe", type=int, default=10)
    ap.add_argument("--steps", type=int, default=12)
    a = ap.parse_args()
    P = importlib.import_module("pumi-pic_b200")
    lib = P.lib()
    lib.pp_ps_set_shuffling(0)
    types = [(np.float64, 3), (np.float64, 3), (np.int32, 1), (np.float64, 3)]
    lo = a.rows // 2
    for split in (1, 0, 1, 0):
        lib.pp_ps_set_rebuild_split_rows(split)
        ppe = np.zeros(a.rows, np.int32)
        ppe[lo:lo + a.occupied] = a.ppe
        ps = P.ParticleStructure(P.capi.PP_PS_SCS, types, ppe)
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        ts = []
        for step in range(a.steps + 3):
            if step == 3:
                torch.cuda.synchronize(); P.api.timing_reset(); P.api.timing_enable(True, 0)
            cap = ps.capacity
            lay = ps.layout()
            se = P.api._tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, ps)
            hop = torch.randint(-30, 31, (cap,), device="cuda", generator=g, dtype=torch.int32)
            stay = torch.rand(cap, device="cuda", generator=g) < 0.5
            ne_new = torch.where(stay, se, (se + hop).clamp(lo, lo + a.occupied - 1)).to(torch.int32).contiguous()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ps.rebuild(ne_new); e1.record(); torch.cuda.synchronize()
            if step >= 3:
                ts.append(e0.elapsed_time(e1))
        table = {k: round(v["avg_ms"], 4) for k, v in P.api.timing_table().items()}
        P.api.timing_enable(False)
        print(json.dumps({"rows": a.rows, "occupied_rows": a.occupied, "particles": int(ps.nptcls), "split_rows": split,
                          "rebuild_ms_median": float(np.median(ts)), "library_phase_avg_ms": table}))
        del ps
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
