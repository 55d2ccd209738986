#!/usr/bin/env python
"""Weak-scaling bench of the FULL PIC step on N GPUs (BASELINE configs[4]); the same record that
bench.py prints as `picstep` (pumi-pic_b200/picstep.py), runnable on its own:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_picstep.py [--cube-per-gpu 55] [--ppe 10] [--steps 10]
"""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube-per-gpu", type=int, default=55)
    ap.add_argument("--ppe", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--push-mult", type=float, default=3.0, help="push distance in units of L/(3*nelems^(1/3))")
    ap.add_argument("--separate-update", action="store_true",
                    help="updatePtclPositions as its own pass instead of a member remap of the rebuild (A/B)")
    ap.add_argument("--overlap-reduce", action="store_true",
                    help="comm-array reduction on its own stream, overlapped with the next step (A/B)")
    ap.add_argument("--timing", action="store_true", help="record the library's phase timers (pp_timing_*)")
    ap.add_argument("--check-steps", type=int, default=2,
                    help="untimed extra steps with the full-size record-multiset check (picstep.check_step)")
    ap.add_argument("--shared-gpu", action="store_true",
                    help="all ranks on cuda:0 (a single-GPU box): hosted communicator without NCCL, gloo for the "
                         "plumbing; the times mean nothing then (the ranks time-slice one GPU), the checks do")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    R = int(os.environ.get("WORLD_SIZE", "1"))
    if a.shared_gpu:
        local = 0
    torch.cuda.set_device(local)
    if R > 1:
        if a.shared_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = importlib.import_module("pumi-pic_b200")
    mod = importlib.import_module("pumi-pic_b200.picstep")
    comm = P.Comm(hosted=True, nccl=False) if (a.shared_gpu and R > 1) else P.Comm()
    r = mod.run_picstep(P, comm, rank, R, a.steps, a.warmup, a.cube_per_gpu, a.ppe, a.push_mult, timing=a.timing,
                        overlap_reduce=a.overlap_reduce, fuse_update=not a.separate_update,
                        full_size_check=a.check_steps)
    if rank == 0:
        print(json.dumps(r))
    if R > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
