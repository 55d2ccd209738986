"""Launches tests/mgpu_worker.py: one rank (serial paths of migrate / reduceCommArray), two ranks that
share one GPU (peer-memory transport between two processes, no NCCL: runs on a single-GPU box), and 2
(or 4) ranks on as many GPUs (NCCL + peer memory over NVLink) when the box has them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(n, p2p=True, only="", shared_gpu=False, hosted=False):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PUMIPIC_P2P="1" if p2p else "0", MGPU_ONLY=only,
               MGPU_EXPECT_P2P="1" if (p2p and n > 1) else "0", MGPU_SHARED_GPU="1" if shared_gpu else "0",
               MGPU_HOSTED="1" if hosted else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(29511 + n),
           os.path.join(HERE, "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_rank_paths():
    _run(1)


def test_two_ranks_on_one_gpu():
    """Two ranks that SHARE cuda:0 (what a single-GPU box can run): communicator from the hosted
    bootstrap without NCCL, migration and comm-array reduction over the peer-memory windows (CUDA IPC
    between the two processes), PIC loop against the serial oracle by particle id, balancer scenario,
    window overflow."""
    try:
        mode = subprocess.run(["nvidia-smi", "--query-gpu=compute_mode", "--format=csv,noheader", "-i", "0"],
                              capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception:
        mode = ""
    if mode and mode != "Default":
        pytest.skip("GPU 0 is in compute mode %r: two processes cannot share it" % mode)
    _run(2, shared_gpu=True)                            # comm arrays, migrate, PIC loop, balancer
    _run(2, shared_gpu=True, only="small_window")


def test_two_ranks_hosted_bootstrap():
    """pp_comm_create_hosted with NCCL: the id travels through the application's all-gather"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _run(2, hosted=True, only="comm_array")
    _run(2, hosted=True, only="migrate")


def test_two_ranks_nccl():
    """all scenarios (incl. the balancer) with the migration over the peer-memory window"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _run(2)


def test_two_ranks_nccl_transport():
    """the migration scenarios again over the NCCL path (AllGather of counts + grouped Send/Recv)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _run(2, p2p=False, only="migrate")
    _run(2, p2p=False, only="pic_loop")


def test_two_ranks_small_window_defers():
    """a window too small for a step's particles: the overflow stays and leaves with a later step"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _run(2, only="small_window")


def test_four_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    _run(4)
