#!/usr/bin/env python3
"""Multi-process check of the fused row-sharded sgemm + all-gather (rb200_sgemm_allgather over CUDA-IPC peer
memory).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1
--master-port 29511 tools/mg_check.py   (gpurun --gpus 2).  Every rank ends with the full C; it is compared with
the NCCL all-gather of the plain stripes (bit-identical) and with a float64 product of sampled rows."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes, sharding
    _capi.init(local)
    _capi.set_stream(torch.cuda.current_stream().cuda_stream)
    f32 = dtypes.float32
    ok = True
    for (M, N, K, mode) in ((2048 * world, 1024, 512, rb.GEMM_TF32), (1000 * world + 3, 520, 4100, rb.GEMM_TF32X3),
                            (4096 * world, 4096, 4096, rb.GEMM_TF32)):
        rng = np.random.default_rng(1234)                          # same A, B on every rank
        A_full = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B_h = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        start, rows = sharding.row_range(M, world, rank)
        A = CudaBuffer(rows * K, f32).from_numpy(A_full[start:start + rows].reshape(-1))
        B = CudaBuffer(K * N, f32).from_numpy(B_h.reshape(-1))
        Cfull = CudaBuffer(M * N, f32, shareable=True)
        pr = sharding.PeerRows(Cfull, M, N)
        blas = rb.MatlibCuda(gemm_mode=mode).blas()
        pr.arrive()
        blas.gemmAllgather(BLAS.NoTrans, BLAS.NoTrans, rows, N, K, 1.0, A, 0, K, B, 0, N, 0.0, Cfull, pr.stripe_offset(), N,
                           pr.peer_ptrs)
        path = _capi.last_gemm_path()
        _capi.sync()
        pr.arrive()
        torch.cuda.synchronize()
        Cfull._dev_dirty = True
        got = Cfull.to_numpy().reshape(M, N)
        # reference: plain stripe + NCCL all-gather
        Cs = CudaBuffer(rows * N, f32)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, rows, N, K, 1.0, A, 0, K, B, 0, N, 0.0, Cs, 0, N)
        _capi.sync()
        want = sharding.gather_rows(sharding.tensor_of(Cs, (rows, N)), M).cpu().numpy()
        same = np.array_equal(got, want)
        idx = rng.integers(0, M, 16)
        ref = A_full[idx].astype(np.float64) @ B_h.astype(np.float64)
        err = np.max(np.abs(got[idx] - ref)) / np.max(np.abs(ref))
        tol = 2e-3 if mode == rb.GEMM_TF32 else 1e-4
        good = same and err < tol
        ok &= good
        print("rank %d: M=%d N=%d K=%d path=%s fused==nccl:%s rel_err=%.2e %s" % (rank, M, N, K, path, same, err,
                                                                                  "ok" if good else "FAIL"), flush=True)
        pr.close()
        del Cfull, A, B, Cs
    # ---- the NVLS form: full C in symmetric memory, stripes multicast from the epilogue (rb200_sgemm_allgather_mc),
    #      and the unicast form over the same symmetric mappings; both against plain stripe + NCCL all-gather
    for (M, N, K, mode) in ((2048 * world, 1024, 512, rb.GEMM_TF32), (1000 * world + 3, 520, 4100, rb.GEMM_TF32X3),
                            (256 * world, 384, 256, rb.GEMM_TF32), (96 * world, 64, 48, rb.GEMM_FP32_SIMT)):
        rng = np.random.default_rng(4321)
        A_full = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B_h = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        blas = rb.MatlibCuda(gemm_mode=mode).blas()
        for allow_mc in (True, False):
            try:
                job = sharding.RowShardedGemm(blas, M, N, K, gather=True, allow_multicast=allow_mc)
            except Exception as e:
                print("rank %d: symmetric memory unavailable: %s: %s" % (rank, type(e).__name__, e), flush=True)
                ok = False
                break
            job.A.from_numpy(A_full[job.start:job.start + job.rows].reshape(-1))
            job.B.from_numpy(B_h.reshape(-1))
            job.C.fill(7.0)
            _capi.sync()
            job.sym.arrive()
            job.run()
            path = _capi.last_gemm_path()
            torch.cuda.synchronize()
            got = job.C.to_numpy().reshape(M, N)
            Cs = CudaBuffer(job.rows * N, f32)
            blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, job.rows, N, K, 1.0, job.A, 0, K, job.B, 0, N, 0.0, Cs, 0, N)
            _capi.sync()
            want = sharding.gather_rows(sharding.tensor_of(Cs, (job.rows, N)), M).cpu().numpy()
            same = np.array_equal(got, want)
            ok &= same
            print("rank %d: symmetric M=%d N=%d K=%d path=%s multicast=%s (ptr %#x) == nccl:%s %s" % (
                rank, M, N, K, path, job.multicast, job.sym.multicast_ptr, same, "ok" if same else "FAIL"), flush=True)
            job.sym.arrive()
            job.close()
            del job, Cs
    # ---- reductions ALONG the sharded axis on NCCL against the oracle on the whole matrix:
    #      NaN mid-column on any rank, at global position 0, at LOCAL position 0 of every later rank, a whole stripe
    #      of NaN, +/-INF ties across rank boundaries
    import oracle
    from rindow_math_matrix_b200 import CudaMath
    for dt, npdt, k in ((dtypes.float32, np.float32, 300), (dtypes.float64, np.float64, 33)):
        Mr = 1000 * world + 5
        rng = np.random.default_rng(99)
        X = (np.round(rng.uniform(-1, 1, (Mr, k)) * 16) / 16).astype(npdt)
        ranges = sharding.all_row_ranges(Mr, world)
        X[5, 3] = np.nan
        X[0, 7] = np.nan
        X[4, 9] = np.inf
        X[:, 10] = -np.inf
        for r, (s0, rr) in enumerate(ranges[1:], start=1):
            X[s0, 2] = np.nan
            X[s0 + 3, 2] = 5.0 + r
            X[s0, 8] = np.nan
            X[s0 + 1, 9] = np.inf
            X[s0 + 7, 4] = np.nan
        X[2, 8] = 70.0
        s_last, r_last = ranges[-1]
        X[s_last:, 6] = np.nan
        X[1, 1] = X[ranges[-1][0] + 2, 1] = 3.0
        start, rows = ranges[rank]
        stripe = CudaBuffer(rows * k, dt).from_numpy(X[start:start + rows].reshape(-1))
        math = rb.MatlibCuda().math()
        sr = sharding.ShardedRows(math, stripe, Mr, k, dt, new_buffer=lambda n, d: CudaBuffer(n, d))
        gi = np.zeros(k, np.int64)
        gm, gs = np.zeros(k), np.zeros(k)
        Xo = oracle.as_buffer(X)
        oracle.reduce_argmax(1, Mr, k, Xo, 0, gi, 0)
        oracle.reduce_max(1, Mr, k, Xo, 0, gm, 0, sticky=True)
        oracle.reduce_sum(1, Mr, k, Xo, 0, gs, 0)
        got_i = sr.reduceArgMax().cpu().numpy()
        got_m = sr.reduceMax().cpu().numpy().astype(np.float64)
        got_s = sr.reduceSum().cpu().numpy().astype(np.float64)
        fin = np.isfinite(gs)
        good_i = np.array_equal(got_i, gi)
        good_m = np.array_equal(got_m, gm, equal_nan=True)
        tol = 1e-5 if dt == dtypes.float32 else 1e-13
        good_s = np.allclose(got_s[fin], gs[fin], rtol=tol, atol=tol * Mr) and np.array_equal(np.isnan(got_s), np.isnan(gs))
        ok &= good_i and good_m and good_s
        print("rank %d: sharded axis-0 reductions [%d,%d] %s over %d ranks (NCCL): argmax exact:%s max exact:%s sum in "
              "tolerance:%s" % (rank, Mr, k, npdt.__name__, world, good_i, good_m, good_s), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
