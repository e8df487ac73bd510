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
        Cfull = CudaBuffer(M * N, f32)
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
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
