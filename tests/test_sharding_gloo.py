"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo.  The per-rank compute is done
with the oracle (test infrastructure); the product code under test is rindow_math_matrix_b200.sharding
(partitioning, gather, merges)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from rindow_math_matrix_b200 import sharding
        rng = np.random.default_rng(42)
        M, N, K = 37, 19, 23                                   # uneven stripes: 19 + 18 rows
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = rng.standard_normal((K, N)).astype(np.float32)
        start, rows = sharding.row_range(M, world, rank)
        # --- row-sharded gemm: local stripe with the oracle, gather with the product code
        c_local = np.zeros(rows * N)
        oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, rows, N, K, 1.0,
                    oracle.as_buffer(A[start:start + rows]), 0, K, oracle.as_buffer(B), 0, N, 0.0, c_local, 0, N)
        full = sharding.gather_rows(torch.from_numpy(c_local.reshape(rows, N)), M).numpy()
        c_ref = np.zeros(M * N)
        oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, M, N, K, 1.0, oracle.as_buffer(A), 0, K,
                    oracle.as_buffer(B), 0, N, 0.0, c_ref, 0, N)
        assert np.array_equal(full, c_ref.reshape(M, N)), "gather_rows (uneven)"
        # even stripes take the all_gather_into_tensor path
        M2 = 36
        s2, r2 = sharding.row_range(M2, world, rank)
        loc = torch.arange(M2 * 3, dtype=torch.float32).view(M2, 3)[s2:s2 + r2].clone()
        assert torch.equal(sharding.gather_rows(loc, M2), torch.arange(M2 * 3, dtype=torch.float32).view(M2, 3))
        # --- reductions along the sharded axis (axis 0)
        X = (np.round(rng.uniform(-1, 1, (M, 11)) * 8) / 8).astype(np.float32)       # many ties
        X[5, 3] = np.nan          # NaN mid-column (rank 0's stripe)
        X[30, 4] = np.nan         # NaN mid-column (rank 1's stripe)
        X[0, 7] = np.nan          # NaN at global position 0
        xs = X[start:start + rows]
        xo = oracle.as_buffer(xs)
        ps, pm, pi = np.zeros(11), np.zeros(11), np.zeros(11, np.int64)
        oracle.reduce_sum(1, rows, 11, xo, 0, ps, 0)
        oracle.reduce_max(1, rows, 11, xo, 0, pm, 0, sticky=True)
        oracle.reduce_argmax(1, rows, 11, xo, 0, pi, 0)
        gs, gm, gi = np.zeros(11), np.zeros(11), np.zeros(11, np.int64)
        Xo = oracle.as_buffer(X)
        oracle.reduce_sum(1, M, 11, Xo, 0, gs, 0)
        oracle.reduce_max(1, M, 11, Xo, 0, gm, 0, sticky=True)
        oracle.reduce_argmax(1, M, 11, Xo, 0, gi, 0)
        got_s = sharding.merge_sum(torch.from_numpy(ps.copy())).numpy()
        ok = ~np.isnan(gs)
        assert np.allclose(got_s[ok], gs[ok], rtol=1e-12, atol=1e-12), "merge_sum"
        got_m = sharding.merge_max(torch.from_numpy(pm.copy())).numpy()
        assert np.array_equal(got_m, gm, equal_nan=True), "merge_max (NaN sticky)"
        # argmax: the stripe records of reduceArgMaxPartial, merged in rank order.  NaNs everywhere that matters:
        # mid-column on either rank, at global position 0, at LOCAL position 0 of rank 1 (a plain NaN globally: the
        # hole round 1 documented), a whole stripe of NaN, +/-INF, ties across the rank boundary
        from oracle_service import OracleBuffer, OracleMath
        s1, _ = sharding.row_range(M, world, 1)
        X[s1, 2] = np.nan                      # NaN at local position 0 of rank 1, real maximum behind it
        X[s1 + 3, 2] = 5.0
        X[s1:, 6] = np.nan                     # rank 1's whole stripe of column 6 is NaN
        X[s1, 8] = np.nan                      # NaN at local 0 of rank 1 AND the global maximum sits on rank 0
        X[2, 8] = 7.0
        X[4, 9] = np.inf
        X[s1 + 1, 9] = np.inf                  # +INF tie across the boundary -> lowest global index
        X[:, 10] = -np.inf                     # all -INF
        X[1, 1] = X[s1 + 2, 1] = 3.0           # finite tie across the boundary
        xs = X[start:start + rows]
        Xo = oracle.as_buffer(X)
        oracle.reduce_sum(1, M, 11, Xo, 0, gs, 0)
        oracle.reduce_max(1, M, 11, Xo, 0, gm, 0, sticky=True)
        oracle.reduce_argmax(1, M, 11, Xo, 0, gi, 0)
        f64 = 13
        stripe = OracleBuffer(rows * 11, f64).from_numpy(xs.reshape(-1))
        om = OracleMath()
        om.sticky_max = True                   # the accelerated-path rule the CUDA kernels follow
        sr = sharding.ShardedRows(om, stripe, M, 11, f64, new_buffer=OracleBuffer,
                                  to_tensor=lambda buf, n, tdt: torch.from_numpy(buf.data[:n]))
        assert (sr.start, sr.rows) == (start, rows)
        got_i = sr.reduceArgMax().numpy()
        assert np.array_equal(got_i, gi), ("ShardedRows.reduceArgMax", got_i, gi)
        assert got_i[7] == 0 and got_i[2] == s1 + 3 and got_i[8] == 2 and got_i[9] == 4 and got_i[10] == 0
        got_m = sr.reduceMax().numpy()
        assert np.array_equal(got_m, gm, equal_nan=True), ("ShardedRows.reduceMax", got_m, gm)
        assert got_m[9] == np.inf                                  # +INF survives the merge (not clamped)
        got_s = sr.reduceSum().numpy()
        ok = np.isfinite(gs)
        assert np.allclose(got_s[ok], gs[ok], rtol=1e-12, atol=1e-12), "ShardedRows.reduceSum"
        # --- PeerRows (fused gemm + gather): handle exchange and stripe addressing, with fake handles
        pr = sharding.PeerRows(None, M, N, export=lambda: bytes([rank + 1]) * 64, open_handle=lambda h: 1000 + h[0])
        assert pr.world == world and pr.rank == rank
        assert (pr.start, pr.rows) == sharding.row_range(M, world, rank)
        assert pr.stripe_offset() == start * N
        assert pr.peer_ptrs == [1000 + r + 1 for r in range(world) if r != rank]     # rank order, self excluded
        assert pr.handles == [bytes([r + 1]) * 64 for r in range(world)]
        pr.arrive()
        pr.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_row_ranges():
    from rindow_math_matrix_b200 import sharding
    assert sharding.all_row_ranges(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sharding.all_row_ranges(8, 8) == [(i, 1) for i in range(8)]
    assert sharding.all_row_ranges(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]
    for total in (1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            rs = sharding.all_row_ranges(total, world)
            assert sum(r for _, r in rs) == total
            assert all(rs[i][0] + rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        sharding.row_range(4, 2, 2)


def test_sharded_path_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
