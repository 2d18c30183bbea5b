"""The ROW-SHARDED eigensolver engine (csrc/symeig.cu: run_symeig_sharded, expand_sharded_kernel, the stop test completed
over the ranks inside rr_kernel) as a host build (tools/emu_engine) on CPU:
  * world = 1: the exchange protocol degenerates to the rank talking to its own region -- eigenpairs against fp64 `eigvalsh`,
    thick restarts (deferred by one matvec), repeated solves on the same regions (epochs);
  * world = 2: two PROCESSES (gloo for the rendezvous only) whose exchange regions are POSIX shared memory mapped into both,
    i.e. the kernels' partial-sum pushes, flags, arrival counters and the peer stores of the new basis block really cross a
    process boundary -- same iteration count and bit-identical eigenvalues on both ranks.
TEST INFRASTRUCTURE (on the GPU box the regions are cudaMalloc'ed and mapped with CUDA IPC: tests/test_gpu_sharded.py)."""
import ctypes
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


class _Patch(object):
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _solve(lib, A_loc, n, k, mb, eps, reg, keep=None):
    from xitorch_b200 import dist as xd
    info = {}
    ev, vec = xd.symeig_row_partitioned(A_loc, n, k, min_eps=eps, info=info, engine="sharded", regions=reg,
                                        max_basis=mb, max_niter=60, restart_keep=keep, gather=False)
    return ev, vec, info


@pytest.mark.parametrize("n,k,mb,dtype,eps", [(256, 4, 64, torch.float64, 1e-6), (256, 4, 32, torch.float64, 1e-6),
                                              (320, 8, 64, torch.float32, 1e-3)])
def test_sharded_engine_world1(emu_lib, monkeypatch, n, k, mb, dtype, eps):
    sys.path.insert(0, HERE)
    import emu_engine_lib
    import oracle
    from xitorch_b200 import dist as xd, _lib
    emu_engine_lib.install(monkeypatch, emu_lib)
    A = oracle.make_herm(n, k, dtype, seed=5)
    pb = emu_lib.xt_symeig_peer_bytes(_lib.dtype_code(dtype), n, k, mb, 1)
    assert pb > 0 and emu_lib.xt_symeig_sharded_workspace_bytes(_lib.dtype_code(dtype), n, k, mb, 1) > 0
    region = torch.zeros(pb, dtype=torch.uint8)
    reg = xd.PeerRegions([region.data_ptr()], pb, 0, 1)
    ref = torch.linalg.eigvalsh(A.double())[:k]
    first = None
    for rep in range(2):                                   # the second solve reuses the regions (next epoch)
        ev, vec, info = _solve(emu_lib, A, n, k, mb, eps, reg)
        assert info["converged"] and info["engine"] == "sharded"
        tol = 1e-9 if dtype == torch.float64 else 1e-5
        assert ((ev.double() - ref).abs() / ref.abs()).max().item() <= tol
        assert (A.double() @ vec.double() - vec.double() * ev.double()).abs().max().item() <= 2 * eps
        if first is None:
            first = (ev.clone(), info["niter"])
        else:
            # run to run the fp64 atomics of the projections land in another order: same answer to rounding (what is
            # bit-identical is the result ACROSS the ranks of one solve, checked by the world-2 test below)
            assert ((first[0].double() - ev.double()).abs() / ref.abs()).max().item() <= tol
            assert abs(first[1] - info["niter"]) <= 1
    assert reg.epoch == 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, so_path):
    from multiprocessing import shared_memory
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    segs = []
    try:
        sys.path.insert(0, HERE)
        import emu_engine_lib
        import oracle
        from xitorch_b200 import dist as xd, _lib
        lib = emu_engine_lib.load(so_path)
        emu_engine_lib.install(_Patch(), lib)
        n, k, mb = 128, 4, 32
        A = oracle.make_herm(n, 2, torch.float64, seed=5)
        lo, hi = xd.shard_range(n, rank, world)
        pb = lib.xt_symeig_peer_bytes(_lib.dtype_code(torch.float64), n, k, mb, world)
        mine = shared_memory.SharedMemory(create=True, size=pb, name="xt_sh_%d_%d" % (port, rank))
        mine.buf[:pb] = bytes(pb)
        segs.append(mine)
        dist.barrier()
        ptrs = []
        for r in range(world):
            seg = mine if r == rank else shared_memory.SharedMemory(name="xt_sh_%d_%d" % (port, r))
            if r != rank:
                segs.append(seg)
            ptrs.append(ctypes.addressof(ctypes.c_char.from_buffer(seg.buf)))
        reg = xd.PeerRegions(ptrs, pb, rank, world)
        out = []
        for rep in range(2):
            info = {}
            ev, vec_loc = xd.symeig_row_partitioned(A[lo:hi].contiguous(), n, k, min_eps=1e-7, info=info, engine="sharded",
                                                    regions=reg, max_basis=mb, max_niter=60, gather=False)
            vecs = [torch.empty_like(vec_loc) for _ in range(world)]
            dist.all_gather(vecs, vec_loc)
            vec = torch.cat(vecs, 0)
            ref = torch.linalg.eigvalsh(A)[:k]
            ok_vals = ((ev - ref).abs() / ref.abs()).max().item() <= 1e-9
            ok_vecs = (A @ vec - vec * ev).abs().max().item() <= 2e-7
            out.append((ok_vals, ok_vecs, info["niter"], info["converged"], ev.tolist()))
        dist.barrier()
        q.put((rank, out))
    finally:
        import gc
        gc.collect()
        for seg in segs:
            try:
                seg.close()
            except BufferError:
                pass
        dist.barrier()
        try:
            segs[0].unlink()
        except Exception:
            pass
        dist.destroy_process_group()


def test_sharded_engine_world2_shared_memory(emu_lib):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, emu_lib.path)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=1500) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        for rep in res[r]:
            assert rep[0] and rep[1] and rep[3], (r, rep)
    for rep in range(2):
        assert res[0][rep][2] == res[1][rep][2]              # the ranks stop at the same iteration ...
        assert res[0][rep][4] == res[1][rep][4]              # ... with bit-identical eigenvalues
