"""CPU-only, 2 gloo ranks: the hosted bootstrap of the communicator (pp_comm_create_hosted,
include/pumipic_b200.h) -- the application's all-gather callback stands where the reference has its
MPI_Comm (support/ViewComm.h).  No GPU here, so no window is ever mapped: what runs is the creation
protocol through the callback (the NCCL id travels through torch.distributed's gloo all-gather), the
size / rank queries, and the error a communicator without NCCL gives for the NCCL-only calls."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, os.path.dirname(HERE))
        import ctypes as C
        import torch                                     # before the library loads NCCL (one libnccl per process)
        import torch.distributed as dist
        pp = importlib.import_module("pumi-pic_b200")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = pp.lib()
        # 1. without NCCL: nothing but the handle is made; the NCCL-only calls say so
        c = pp.Comm(world, rank, hosted=True, nccl=False)
        assert lib.pp_comm_size(c.h) == world and lib.pp_comm_rank(c.h) == rank
        assert not c.p2p_active
        buf = (C.c_int32 * (2 * world))()
        rc = lib.pp_comm_alltoall(c.h, buf, buf, 1, pp.capi.PP_INT32, None)
        assert rc != 0 and "no NCCL transport" in lib.pp_last_error().decode(), lib.pp_last_error().decode()
        rc = lib.pp_comm_send(c.h, buf, 1, pp.capi.PP_INT32, (rank + 1) % world, None)
        assert rc != 0 and "no NCCL transport" in lib.pp_last_error().decode()
        # 2. the callback itself: bytes of every rank, in rank order
        seen = []
        def allgather(_ctx, send, recv, nbytes):
            parts = [None] * world
            dist.all_gather_object(parts, C.string_at(send, nbytes))
            seen.append(parts)
            C.memmove(recv, b"".join(parts), nbytes * world)
            return 0
        cb = pp.capi.HOST_ALLGATHER_FN(allgather)
        h = C.c_void_p()
        # with NCCL requested the id goes through the callback; the communicator itself cannot come
        # up without a GPU, and both ranks get the error instead of a hang
        rc = lib.pp_comm_create_hosted(world, rank, C.cast(cb, C.c_void_p), None, 1, C.byref(h))
        assert len(seen) == 1 and len(seen[0]) == world and all(len(p) == 128 for p in seen[0])
        assert seen[0][0] != bytes(128) and all(p == bytes(128) for p in seen[0][1:]), "rank 0 provides the id"
        assert rc != 0 and not h.value, "no GPU: the NCCL communicator must fail, not hang"
        # 3. argument checks
        rc = lib.pp_comm_create_hosted(world, rank, None, None, 0, C.byref(h))
        assert rc != 0 and "all-gather callback is required" in lib.pp_last_error().decode()
        rc = lib.pp_comm_create_hosted(1, 0, None, None, 0, C.byref(h))      # one rank needs no callback
        assert rc == 0 and lib.pp_comm_size(h) == 1
        lib.pp_comm_destroy(h)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "fail: %r\n%s" % (e, traceback.format_exc())))


def test_hosted_bootstrap_on_two_gloo_ranks():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29960 + os.getpid() % 30
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
