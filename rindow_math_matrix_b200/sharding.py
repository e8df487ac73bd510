"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

SURVEY.md section 8(e): the path shards by ROWS.
  * sgemm            : rank g owns rows [start_g, start_g + rows_g) of A and C, B is replicated; each
                       rank runs the single-GPU kernel on its stripe.  No data-path collective unless
                       the caller wants the full C everywhere -> gather_rows() (NCCL all-gather).
  * add / multiply / softmax / reductions along a non-sharded axis / im2col2d: independent units,
                       no collective at all.
  * reductions ALONG the sharded axis (axis 0 of row-sharded data): local partials [k] per rank, then
                       one small exchange -- merge_sum / merge_max / merge_argmax below.  max and argmax
                       gather the per-rank partials and merge them locally in rank order, so NaN
                       stickiness and the lowest-global-index tie rule of the single-GPU kernels
                       (PhpMath.php:95-130) are preserved exactly and deterministically.

Everything here works on torch tensors of any device, so the same code runs under NCCL on B200s and
under gloo on CPU (tests/test_sharding_gloo.py, world_size 2).  No compute happens in this module.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def row_range(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row stripe of `rank`: the first (total_rows % world) ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total_rows, world)
    rows = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, rows


def all_row_ranges(total_rows: int, world: int) -> List[Tuple[int, int]]:
    return [row_range(total_rows, world, r) for r in range(world)]


def tensor_of(buffer, shape=None, offset: int = 0) -> torch.Tensor:
    """Zero-copy torch view of a CudaBuffer (through __cuda_array_interface__), for NCCL."""
    t = torch.as_tensor(buffer, device="cuda")
    if offset:
        t = t[offset:]
    if shape is not None:
        n = 1
        for s in shape:
            n *= int(s)
        t = t[:n].view(*shape)
    return t


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def gather_rows(local: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row stripes [rows_g, ...] into the full [total_rows, ...] tensor on every rank."""
    world = _world(group)
    if world == 1:
        return local
    ranges = all_row_ranges(total_rows, world)
    tail = tuple(local.shape[1:])
    if len({r for _, r in ranges}) == 1:
        out = torch.empty((total_rows,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven stripes: pad to the largest, gather, trim
    max_rows = max(r for _, r in ranges)
    padded = torch.zeros((max_rows,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:r] for p, (_, r) in zip(parts, ranges)], dim=0)


class PeerRows:
    """A full [total_rows, cols] float32 matrix replicated on every rank whose row stripes are produced by their
    owners and pushed to all peers over NVLink from inside the producing kernel (rb200_sgemm_allgather) -- the
    fused form of `gemm` + `gather_rows`.  One process per GPU: the full-C allocations are exchanged as CUDA IPC
    handles through the process group (host plumbing only) and mapped with rb200_ipc_open; `full_buffer` must be a
    `CudaBuffer(..., shareable=True)` (plain cudaMalloc memory: pool memory cannot be exported).  SymmetricRows below
    is the NVLS / symmetric-memory form the benchmark uses.

    exchange() is injectable so the handle plumbing runs under gloo on CPU (tests/test_sharding_gloo.py)."""

    def __init__(self, full_buffer, total_rows: int, cols: int, group=None, export=None, open_handle=None):
        from . import _capi
        import ctypes
        self.buffer, self.total_rows, self.cols, self.group = full_buffer, int(total_rows), int(cols), group
        self.world = _world(group)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.start, self.rows = row_range(self.total_rows, self.world, self.rank)
        self._opened = []
        self._flag = None
        self.peer_ptrs: List[int] = []
        if self.world == 1:
            return
        lib = _capi.load() if export is None else None

        def _export():
            h = (ctypes.c_uint8 * 64)()
            _capi.check(lib.rb200_ipc_export(full_buffer.raw_pointers()[0], h))
            return bytes(h)

        def _open(handle: bytes) -> int:
            p = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
            _capi.check(lib.rb200_ipc_open(hb, ctypes.byref(p)))
            self._opened.append(p.value)
            return p.value

        handles = [None] * self.world
        dist.all_gather_object(handles, (export or _export)(), group=group)
        self.handles = handles
        self.peer_ptrs = [(open_handle or _open)(h) for r, h in enumerate(handles) if r != self.rank]

    def stripe_offset(self) -> int:
        """element offset of this rank's stripe inside the full matrix"""
        return self.start * self.cols

    def arrive(self) -> None:
        """Order the ranks after a step: every stripe has landed everywhere once this returns on the stream
        (a process-group barrier; under NCCL a tiny all-reduce on the current stream)."""
        if self.world > 1:
            if self._flag is None:
                on_gpu = dist.get_backend(self.group) == "nccl"
                self._flag = torch.zeros(1, dtype=torch.int32, device="cuda" if on_gpu else "cpu")
            dist.all_reduce(self._flag, group=self.group)         # stream-ordered under NCCL: no host sync

    def close(self) -> None:
        if self._opened:
            from . import _capi
            lib = _capi.load()
            for p in self._opened:
                lib.rb200_ipc_close(p)
            self._opened = []


class SymmetricRows:
    """The full [total_rows, cols] float32 matrix in SYMMETRIC memory: one allocation per rank, every rank's copy
    mapped into every process and all of them bound to one NVLS multicast object
    (torch.distributed._symmetric_memory: cuMemCreate + cuMulticastCreate / cuMulticastBindMem, handles exchanged
    through the process group -- plumbing only).  `multicast_ptr` is the alias a multimem.st writes through: the
    NVSwitch replicates the store into all copies, so the producing GEMM's epilogue (rb200_sgemm_allgather_mc) sends
    a stripe over its NVLink ports ONCE instead of once per peer.  Without multicast support (`multicast_ptr == 0`)
    the peers' mappings (`peer_ptrs`) serve the unicast form, as PeerRows does over CUDA IPC.
    arrive() is the symmetric-memory barrier (a signal-pad kernel on the current stream, no host sync)."""

    def __init__(self, total_rows: int, cols: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import dtypes
        from .cuda_buffer import CudaBuffer
        self.total_rows, self.cols = int(total_rows), int(cols)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.start, self.rows = row_range(self.total_rows, self.world, self.rank)
        n = self.total_rows * self.cols
        self.tensor = symm_mem.empty(n, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.handle = symm_mem.rendezvous(self.tensor, self.group)
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        ptrs = list(self.handle.buffer_ptrs)
        self.peer_ptrs = [int(p) for r, p in enumerate(ptrs) if r != self.rank]
        self.buffer = CudaBuffer.adopt(self.tensor.data_ptr(), n, dtypes.float32, owner=self.tensor)

    def stripe_offset(self) -> int:
        return self.start * self.cols

    def arrive(self) -> None:
        self.handle.barrier(channel=0)

    def close(self) -> None:
        self.buffer = None
        self.handle = None
        self.tensor = None


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process to the CPU cores NVML reports as local to GPU `local_rank` (its NUMA node) BEFORE any pinned
    host memory is allocated: cudaHostAlloc places pages on the allocating thread's node, and a rank whose mirrors sit
    on the far socket pays the inter-socket hop on every h2d / d2h.  Returns {"cpus": n, "numa": id or None}; a no-op
    ({"cpus": 0}) when NVML or the affinity call is unavailable (single-socket boxes lose nothing)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"cpus": 0, "numa": None}
        os.sched_setaffinity(0, cpus)
        numa = None
        try:
            numa = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        return {"cpus": len(cpus), "numa": numa}
    except Exception:
        return {"cpus": 0, "numa": None}


class RowShardedGemm:
    """C = alpha * A . B (+ beta * C) for float32 with A and C ROW-SHARDED over the ranks of `group` and B replicated
    (SURVEY.md section 8(e) row 1): rank g owns rows [start, start + rows) of the [M, K] matrix A and of the [M, N]
    result, and runs the single-GPU kernel on its stripe.  Two ways to finish:
      * gather=False -- C stays sharded (no collective at all);
      * gather=True  -- every rank ends with the FULL C: the stripes are pushed into all ranks' copies from the GEMM
        epilogue while the remaining tiles are still being multiplied (NVLS multicast stores when the NVSwitch offers
        them, unicast peer stores otherwise), then one barrier.
    run() works on device-resident operands.  run_from_host() is the host-operand form of the same call: every rank
    uploads ITS stripe of A and ITS 1/world slice of B from pinned memory, the slices of B are all-gathered over
    NVLink (so B crosses each PCIe link once in total, not once per rank), the stripe is multiplied by the
    transfer-pipelined rb200_sgemm_streamed and the C stripe comes back to the host while later row blocks multiply."""

    def __init__(self, blas, M: int, N: int, K: int, group=None, gather: bool = True, allow_multicast: bool = True,
                 operands=None):
        from . import dtypes
        from .cuda_buffer import CudaBuffer
        self.blas, self.M, self.N, self.K, self.group = blas, int(M), int(N), int(K), group
        self.world = _world(group)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.start, self.rows = row_range(self.M, self.world, self.rank)
        # this rank's slice of B's rows for the host-operand form; when K does not divide evenly every rank uploads all
        # of B itself (the in-place all-gather needs equal slices)
        if self.K % self.world == 0:
            self.bstart, self.brows = row_range(self.K, self.world, self.rank)
        else:
            self.bstart, self.brows = 0, self.K
        if self.rows < 1:
            raise ValueError("every rank needs at least one row of A")
        f32 = dtypes.float32
        if operands is not None:                 # share another job's resident A stripe and B
            self.A, self.B = operands
        else:
            self.A = CudaBuffer(self.rows * self.K, f32, zero=False)
            self.B = CudaBuffer(self.K * self.N, f32, zero=False)
        self.gather = bool(gather) and self.world > 1
        self.sym = None
        self.multicast = False
        if self.gather:
            self.sym = SymmetricRows(self.M, self.N, group)
            self.C, self.offC = self.sym.buffer, self.sym.stripe_offset()
            self.multicast = bool(allow_multicast and self.sym.multicast_ptr)
        else:
            self.C, self.offC = CudaBuffer(self.rows * self.N, f32), 0
        self._host = None

    # ---- device-resident ------------------------------------------------------------------------------
    def run(self, alpha: float = 1.0, beta: float = 0.0) -> None:
        from .cuda_blas import BLAS
        if not self.gather:
            self.blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, self.rows, self.N, self.K, alpha, self.A, 0, self.K,
                           self.B, 0, self.N, beta, self.C, self.offC, self.N)
            return
        self.blas.gemmAllgather(BLAS.NoTrans, BLAS.NoTrans, self.rows, self.N, self.K, alpha, self.A, 0, self.K,
                                self.B, 0, self.N, beta, self.C, self.offC, self.N, self.sym.peer_ptrs,
                                multicast_ptr=self.sym.multicast_ptr if self.multicast else None)
        self.sym.arrive()

    # ---- host operands ----------------------------------------------------------------------------------
    def host_buffers(self):
        """(A stripe [rows, K], B slice [brows, N], C stripe [rows, N]) as numpy views of pinned host memory the caller
        fills / reads; created on first use on the calling thread's NUMA node (bind_to_gpu_numa first)."""
        if self._host is None:
            from . import dtypes
            from .cuda_buffer import CudaBuffer
            f32 = dtypes.float32
            self._Bslice = CudaBuffer(max(1, self.brows * self.N), f32, zero=False)    # staging mirror of B's slice
            self._Cstripe = self.C if not self.gather else CudaBuffer(self.rows * self.N, f32)
            self._host = (self.A.host().reshape(self.rows, self.K), self._Bslice.host().reshape(-1)[: self.brows * self.N]
                          .reshape(self.brows, self.N), self._Cstripe.host().reshape(self.rows, self.N))
        return self._host

    def run_from_host(self, alpha: float = 1.0) -> None:
        """One step with this step's inputs in the pinned host buffers and the C stripe read back into its host buffer
        (returns when the stripe is on the host)."""
        from . import _capi
        from .cuda_blas import BLAS
        self.host_buffers()
        lib = _capi.load()
        # B: own slice up (asynchronously on the library stream), then the NVLink all-gather in place
        dB, _ = self.B.raw_pointers()
        _, hS = self._Bslice.raw_pointers()
        self._Bslice._wait_upload()
        if self.brows:
            _capi.check(lib.rb200_buffer_h2d_async(dB, self.bstart * self.N * 4, hS, self.brows * self.N * 4))
        if self.world > 1 and self.brows < self.K:
            tB = tensor_of(self.B)
            dist.all_gather_into_tensor(tB, tB[self.bstart * self.N:(self.bstart + self.brows) * self.N], group=self.group)
        self.A.mark_host_written()
        previous = self.blas.host_result_prefetch
        self.blas.setHostResultPrefetch(True)
        try:
            self.blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, self.rows, self.N, self.K, alpha, self.A, 0, self.K,
                           self.B, 0, self.N, 0.0, self._Cstripe, 0, self.N)
        finally:
            self.blas.setHostResultPrefetch(previous)
        self._Cstripe.host()                      # current: the streamed call already brought it back

    def close(self) -> None:
        if self.sym is not None:
            self.sym.close()
            self.sym = None


def merge_sum(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of per-rank partial sums (all-reduce)."""
    if _world(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def _gather_stack(t: torch.Tensor, group=None) -> torch.Tensor:
    world = _world(group)
    if world == 1:
        return t.unsqueeze(0)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous(), group=group)
    return torch.stack(parts, dim=0)


def merge_max(partial: torch.Tensor, group=None) -> torch.Tensor:
    """NaN-sticky maximum over ranks of per-rank partial maxima."""
    stack = _gather_stack(partial, group)                  # [world, k]
    has_nan = torch.isnan(stack).any(dim=0)
    # (not nan_to_num: it would also clamp +/-INF to the largest finite value)
    mx = torch.where(torch.isnan(stack), torch.full_like(stack, float("-inf")), stack).max(dim=0).values
    return torch.where(has_nan, torch.full_like(mx, float("nan")), mx)


def merge_argmax(values: torch.Tensor, indices: torch.Tensor, row_offset: int, group=None) -> torch.Tensor:
    """Global arg-max along the sharded axis from per-rank (value, stripe-local index) records as
    rb200_reduce_argmax_partial / CudaMath.reduceArgMaxPartial writes them: the rank that owns global position 0
    follows math_imax (PhpMath.php:114-130) with its never-displaced NaN at position 0 reported as +INF; every other
    rank skips NaNs (`$value > $max` is false for a NaN wherever it sits) and reports (-INF, INT32_MAX) for an all-NaN
    stripe.  Merged in rank order with a strict `>`: first strict maximum wins, ties -> lowest global index, a NaN
    only wins at GLOBAL position 0 -- bit-identical to the single-buffer result, NaNs included."""
    gidx = indices.to(torch.int64) + int(row_offset)
    vals = _gather_stack(values, group)                    # [world, k]
    idxs = _gather_stack(gidx, group)
    best_v, best_i = vals[0].clone(), idxs[0].clone()
    for r in range(1, vals.shape[0]):
        better = vals[r] > best_v                          # strict: ties stay with the lower rank = lower index
        best_v = torch.where(better, vals[r], best_v)
        best_i = torch.where(better, idxs[r], best_i)
    return best_i


def split_records(records: torch.Tensor, outputs: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values float64 [outputs], indices int64 [outputs]) views of the int64 record buffer
    reduceArgMaxPartial fills (two words per output: the bits of the double value, the index)."""
    rec = records[: 2 * outputs].view(outputs, 2)
    return rec[:, 0].contiguous().view(torch.float64), rec[:, 1].contiguous()


class ShardedRows:
    """A [total_rows, k] matrix row-sharded over the ranks of `group`: this rank holds rows
    [start, start + rows) in a driver Buffer.  reduce*(axis=0) run the single-GPU kernel on the stripe
    (math.reduceSum / reduceMax / reduceArgMaxPartial with (m, n, k) = (1, rows, k)) and finish with ONE small
    exchange of [k]-sized partials: all-reduce for the sum, gathered partials
    merged in rank order for max / argmax.  Every rank returns the full [k] result as a torch tensor on the
    partials' device.  `to_tensor(buffer, n, torch_dtype)` views a driver buffer as a tensor (CUDA: zero-copy
    through __cuda_array_interface__; the gloo tests pass a numpy-backed stand-in)."""

    def __init__(self, math, buffer, total_rows: int, k: int, dtype: int, new_buffer, to_tensor=None, group=None):
        self.math, self.buffer, self.total_rows, self.k, self.dtype = math, buffer, int(total_rows), int(k), dtype
        self.group, self.new_buffer = group, new_buffer
        self.world = _world(group)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.start, self.rows = row_range(self.total_rows, self.world, self.rank)
        self.to_tensor = to_tensor or (lambda buf, n, tdt: tensor_of(buf)[:n])
        if self.rows < 1:
            raise ValueError("every rank needs at least one row of the sharded axis")

    def reduceSum(self) -> torch.Tensor:
        part = self.new_buffer(self.k, self.dtype)
        self.math.reduceSum(1, self.rows, self.k, self.buffer, 0, part, 0)
        return merge_sum(self.to_tensor(part, self.k, None), self.group)

    def reduceMax(self) -> torch.Tensor:
        part = self.new_buffer(self.k, self.dtype)
        self.math.reduceMax(1, self.rows, self.k, self.buffer, 0, part, 0)
        return merge_max(self.to_tensor(part, self.k, None), self.group)

    def reduceArgMax(self) -> torch.Tensor:
        from . import dtypes
        rec = self.new_buffer(2 * self.k, dtypes.int64)
        self.math.reduceArgMaxPartial(1, self.rows, self.k, self.buffer, 0, self.start == 0, rec, 0)
        vals, idxs = split_records(self.to_tensor(rec, 2 * self.k, torch.int64), self.k)
        return merge_argmax(vals, idxs, self.start, self.group)
