"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the
row-partitioned operator with its one all-gather per application."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xitorch_b200 import dist as xd
        import oracle
        torch.manual_seed(0)
        n, k = 64, 4
        A = oracle.make_herm(n, 2, torch.float64, seed=3)
        X = torch.randn(n, k, dtype=torch.float64)
        lo, hi = xd.shard_range(n, rank, world)
        op = xd.RowPartitionedOperator(A[lo:hi].contiguous(), n)
        Y = op.mm(X)
        ok_mm = torch.allclose(Y, A @ X, rtol=1e-12, atol=1e-12)
        yv = op.mv(X[:, 0])
        ok_mv = torch.allclose(yv, A @ X[:, 0], rtol=1e-12, atol=1e-12)
        # the operator is usable by any method that only needs mm: exact reference algorithm via the oracle-free
        # exacteig path needs fullmatrix -> also one collective per column block
        full = op.fullmatrix()
        ok_full = torch.allclose(full, A)
        # batch sharding: contiguous, disjoint, covering
        B = torch.arange(10.0).reshape(5, 2)
        mine = xd.shard_batch(B)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]]))
        ok_shard = sum(int(s.item()) for s in sizes) == 5 and mine.shape[0] in (2, 3)
        # a Krylov method on a CPU row-partitioned operator must refuse (no CPU fallback)
        try:
            xd.symeig_row_partitioned(A[lo:hi].contiguous(), n, 2)
            ok_refuse = False
        except RuntimeError as e:
            ok_refuse = "CUDA" in str(e)
        q.put((rank, ok_mm, ok_mv, ok_full, ok_shard, ok_refuse, op.napply))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_operator_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:6]), r
        assert r[6] >= 3          # mm, mv, fullmatrix each applied the operator (one all-gather each)


def _worker_engine(rank, world, port, q):
    """the row-partitioned eigensolver entry with the CUDA library replaced by the stand-in: checks the all-gather hook
    the host installs (pointer arithmetic inside the workspace, in-place gather of `world` chunks) over gloo"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import standin_engine
        from xitorch_b200 import dist as xd
        import oracle

        class _Patch(object):
            def setattr(self, obj, name, value):
                setattr(obj, name, value)

        eng = standin_engine.install(_Patch())
        n, k = 48, 4
        A = oracle.make_herm(n, 2, torch.float64, seed=5)
        lo, hi = xd.shard_range(n, rank, world)
        info = {}
        evals, evecs = xd.symeig_row_partitioned(A[lo:hi].contiguous(), n, k, info=info)
        ref = torch.linalg.eigvalsh(A)[:k]
        ok_vals = torch.allclose(evals, ref, atol=1e-10)
        ok_vecs = torch.allclose(A @ evecs, evecs * evals, atol=1e-9)
        eig_world = eng.log[-1]["world"]
        # batch-sharded independent systems (BASELINE config 3): no data-path collective, reduced bookkeeping only
        g = torch.Generator().manual_seed(11)
        nbt, m = 5, 10
        Ab = torch.eye(m, dtype=torch.float64) + 0.3 * torch.randn(nbt, m, m, generator=g, dtype=torch.float64) / m ** 0.5
        Bb = torch.randn(nbt, m, 1, generator=g, dtype=torch.float64)
        x_local, sinfo = xd.solve_batch_sharded(xd.shard_batch(Ab), xd.shard_batch(Bb), method="bicgstab")
        ok_solve = torch.allclose(xd.shard_batch(Ab) @ x_local, xd.shard_batch(Bb), atol=1e-10)
        ok_info = sinfo["all_converged"] and sinfo["niter_max"] >= 1 and eng.log[-1]["nbatch"] in (2, 3)
        q.put((rank, ok_vals and ok_solve and ok_info, ok_vecs, eig_world, tuple(evecs.shape)))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_symeig_allgather_hook_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_engine, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2], r
        assert r[3] == world and r[4] == (48, 4)


def _worker_emulated(rank, world, port, q, so_path):
    """the row-partitioned path of the REAL engine (host build of csrc/symeig.cu, tools/emu_engine) on two gloo ranks:
    local row block times the basis block, flag row packed behind it, one in-place all-gather per application, the
    gathered image unpacked and the stop decision taken from the gathered flags -- on every rank identically"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import emu_engine_lib
        from xitorch_b200 import dist as xd
        import oracle

        class _Patch(object):
            def setattr(self, obj, name, value):
                setattr(obj, name, value)

        emu_engine_lib.install(_Patch(), emu_engine_lib.load(so_path))
        n, k = 64, 4
        A = oracle.make_herm(n, 2, torch.float64, seed=5)
        lo, hi = xd.shard_range(n, rank, world)
        info = {}
        evals, evecs = xd.symeig_row_partitioned(A[lo:hi].contiguous(), n, k, min_eps=1e-8, info=info)
        ref = torch.linalg.eigvalsh(A)[:k]
        ok_vals = ((evals - ref).abs() / ref.abs()).max().item() <= 1e-9
        ok_vecs = (A @ evecs - evecs * evals).abs().max().item() <= 2e-7
        q.put((rank, ok_vals, ok_vecs, info["niter"], info["converged"], evals.tolist()))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_engine_emulated_world2(emu_lib):
    lib = emu_lib
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_emulated, args=(r, world, port, q, lib.path)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2] and r[4], r
    assert res[0][3] == res[1][3]                    # both ranks stopped at the same iteration
    assert res[0][5] == res[1][5]                    # ... with bit-identical eigenvalues (replicated small algebra)


def test_shard_range_properties():
    from xitorch_b200.dist import shard_range
    for n in (1, 7, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_row_partition_needs_divisible_n():
    from xitorch_b200.dist import RowPartitionedOperator
    A = torch.zeros(5, 10)
    with pytest.raises(RuntimeError):
        RowPartitionedOperator(A, 11)
