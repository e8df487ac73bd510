#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 driver for rindow-math-matrix.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference's CPU algorithm
  torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU

Metric (BASELINE.json): "sgemm TFLOP/s; add/reduceSum HBM GB/s at 1/2/4/8 B200 vs PHP CPU path".
The step is ONE float32 GEMM C = A*B (NoTrans/NoTrans, alpha=1, beta=0) of M=N=K=16384 (config C5, the largest size
of the sweep; 3 GiB of operands) issued through the reference-facing plugin interface (service.blas().gemm(...) with
the PhpBlas signature) down the C ABI.  With N GPUs the SAME problem is row-sharded (strong scaling): rank g owns
rows [g*16384/N, (g+1)*16384/N) of A and C, B is replicated, and every rank ends the step holding the full C -- the
stripes are stored into all ranks' copies from the GEMM epilogue over NVLink (NVLS multicast when available) while
the remaining tiles multiply, followed by one barrier ("with_gather.compute_only" is the same step with C left
sharded, "with_gather.nccl" the unfused library all-gather for context).  `value` times device-resident operands;
`e2e` times the same job with HOST buffers (per rank: its A stripe and its 1/N slice of B from pinned memory -> h2d,
B's slices all-gathered over NVLink, transfer-pipelined multiply, C stripe d2h) inside the timed region.  The
HBM-bound configs (C2/C3/C4: add, multiply, reduceSum/Max/ArgMax, softmax, im2col2d) are reported in "hbm_ops" of
the same JSON line with their own roofline fractions and CPU baselines; other GEMM modes (3xTF32 parity mode, FFMA,
float64) in "gemm_modes"; the C5 size sweep in "sgemm_sweep".

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sgemm TFLOP/s"
UNIT = "TFLOP/s"


# Only the JSON line may reach stdout: NCCL / CUDA libraries print banners to fd 1 from C, so fd 1 is
# pointed at stderr for the whole run and the line is written to a saved copy of the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        # fallback stated by /opt/skills/guides/B200_PROFILING.md
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/traffic.json: "<kernel>@<n>" -> bytes); None when no capture exists for this kernel/size."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("%s@%d" % (kernel, n))
    except Exception:
        return None


def workload_config(n, world, mode_name):
    return {
        "workload": "sgemm M=N=K=%d (BASELINE config C5, the largest size of the sweep), NoTrans/NoTrans, alpha=1, "
                    "beta=0; ONE problem row-sharded over the GPUs" % n,
        "gemm_mode": mode_name,
        "tolerance": {"tf32": "max|err| < 2e-3*max|C| (1xTF32; outside the reference isclose rtol 1e-4 by design)",
                      "tf32x3": "max|err| < 1e-4*max|C| (reference isclose bound)",
                      "fp32_simt": "max|err| < 1e-4*max|C| (reference isclose bound)"}[mode_name],
        "global_shape": [n, n, n],
        "parallelism": ("A and C row-sharded over %d GPU(s) (%d rows each), B replicated; every rank ends with the full C "
                        "(stripes pushed from the GEMM epilogue over NVLink, one barrier)" % (world, n // world))
                       if world > 1 else "1 GPU",
        "l2": "inputs exceed L2: A stripe %d MiB + B %d MiB per GPU vs 126 MB L2" % (n * n * 4 // world >> 20, n * n * 4 >> 20),
    }


# -------------------------------------------------------------------------------------------------
# clocks: NVML sampler thread (nvidia-smi would give <1 sample in a 50 ms region)
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, local_rank):
        self.samples, self.reasons, self.ok = [], set(), False
        self.active = False             # only the timed regions are sampled (set by timed())
        self._stop = threading.Event()
        self.sm_max = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if not self.active:
                time.sleep(0.0005)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._stop.set()
            self.t.join()

    def nvlink_bytes(self):
        """(tx, rx) payload bytes this GPU has moved over all its NVLink ports since boot (NVML field values
        NVLINK_THROUGHPUT_DATA_TX / RX, KiB, scope = all links); None when the driver does not expose them."""
        if not self.ok:
            return None
        try:
            nv = self.nv
            ids = [nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX]
            try:
                vals = nv.nvmlDeviceGetFieldValues(self.h, [(i, 0xFFFFFFFF) for i in ids])   # scopeId = all links
            except Exception:
                vals = nv.nvmlDeviceGetFieldValues(self.h, ids)
            out = []
            for v in vals:
                if v.nvmlReturn != 0:
                    raise RuntimeError("field value not available")
                out.append(int(v.value.ullVal) * 1024)
            return tuple(out)
        except Exception:
            pass
        # fallback: `nvidia-smi nvlink -gt d` (per-link "Data Tx / Rx: N KiB" lines of this GPU)
        try:
            import re
            import subprocess
            idx = self.nv.nvmlDeviceGetIndex(self.h)
            txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(idx)], capture_output=True, text=True,
                                 timeout=20).stdout
            tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
            rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
            if tx == 0 and rx == 0:
                return None
            return (tx * 1024, rx * 1024)
        except Exception:
            return None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (C restatement of PhpBlas::gemm) on the host cores
# -------------------------------------------------------------------------------------------------
def reference_gemm_sample(n, rows_total, threads, A64=None, B64=None):
    """Compute `rows_total` rows of C = A*B with the reference loop order (oracle.gemm_rows), spread
    over `threads` host threads (ctypes releases the GIL).  Returns (seconds, rows, C_rows)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    if A64 is None:
        lg = int(np.log2(n))
        A64 = oracle.as_buffer(np.random.default_rng(5000 + lg).uniform(-1, 1, n * n).astype(np.float32))
        B64 = oracle.as_buffer(np.random.default_rng(5100 + lg).uniform(-1, 1, n * n).astype(np.float32))
    C = np.zeros(rows_total * n)
    rows = list(range(rows_total))
    chunks = [rows[i::threads] for i in range(threads)]

    def work(chunk):
        for r in chunk:
            oracle.gemm_rows(r, 1, n, n, 1.0, A64, n, B64, n, 0.0, C, n)     # row r of A -> row r of the sample

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, chunks))
    return time.perf_counter() - t0, rows_total, C.reshape(rows_total, n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    import oracle
    n = args.n
    cores = os.cpu_count() or 1
    lg = int(np.log2(n))
    A64 = oracle.as_buffer(np.random.default_rng(5000 + lg).uniform(-1, 1, n * n).astype(np.float32))
    B64 = oracle.as_buffer(np.random.default_rng(5100 + lg).uniform(-1, 1, n * n).astype(np.float32))
    # calibrate one row per thread, then size the per-step sample so the run stays within ~2.5 minutes
    t_cal, _, _ = reference_gemm_sample(n, cores, cores, A64, B64)
    budget = 150.0
    total_steps = args.steps + args.warmup
    rows_per_step = cores
    if t_cal * total_steps > budget:
        rows_per_step = max(1, int(cores * budget / (t_cal * total_steps)))
    threads = min(cores, rows_per_step)
    for _ in range(args.warmup):
        reference_gemm_sample(n, rows_per_step, threads, A64, B64)
    t = 0.0
    for _ in range(args.steps):
        dt, _, _ = reference_gemm_sample(n, rows_per_step, threads, A64, B64)
        t += dt
    flops = 2.0 * rows_per_step * n * n * args.steps
    value = flops / t / 1e12
    sample = ("%d of %d rows of C per step (N=K=%d), reference loop order m,n,k in double "
              "(C restatement of PhpBlas::gemm; the PHP interpreter is not installable here and would be "
              "much slower), %d thread(s)" % (rows_per_step, n, n, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(n, world, args.mode),
        "arithmetic": "reference (double accumulation in the reference's loop order)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# -------------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        log("bench.py --gpus %d without torchrun: running rank 0 only; launch with torch.distributed.run" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 driver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    from rindow_math_matrix_b200 import sharding
    numa = sharding.bind_to_gpu_numa(local_rank)       # before any pinned allocation: mirrors land on the GPU's node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes

    _capi.init(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    _capi.set_stream(stream.cuda_stream)              # kernels + events on the same stream
    svc = rb.MatlibCuda()
    assert svc.serviceLevel() >= rb.Service.LV_ADVANCED
    blas, math = svc.blas(), svc.math()
    peaks = measured_peaks()
    mode = {"tf32": rb.GEMM_TF32, "tf32x3": rb.GEMM_TF32X3, "fp32_simt": rb.GEMM_FP32_SIMT}[args.mode]
    sampler = ClockSampler(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    per_rank = {}

    def timed(fn, steps, warmup, tag=None):
        """W untimed calls, then exactly `steps` calls between two events on the launching stream,
        barrier + synchronize on both sides, max over ranks.  Returns total ms.  `tag` also records every
        rank's own time (per_rank[tag], ms per step, rank order)."""
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sampler.active = True
            e0.record(stream)
            for i in range(steps):
                fn()
            e1.record(stream)
            barrier()
            sampler.active = False
        mine = e0.elapsed_time(e1)
        if tag is not None:
            if world > 1:
                t = torch.zeros(world, dtype=torch.float64, device="cuda")
                t[rank] = mine / steps
                dist.all_reduce(t)
                per_rank[tag] = [float(x) for x in t.tolist()]
            else:
                per_rank[tag] = [mine / steps]
        return max_over_ranks(mine)

    # ---- the problem: ONE sgemm n^3, A and C row-sharded, B replicated; operands resident in HBM ---------------
    n = args.n
    if n % world:
        raise SystemExit("bench.py: --size must be divisible by the number of GPUs")
    rows = n // world
    lg = int(np.log2(n))
    # rank r's stripe of A is drawn from its own seed (the global A is the concatenation), B from one seed everywhere
    Ah = np.random.default_rng(5000 + lg + 1000 * rank).uniform(-1, 1, rows * n).astype(np.float32)
    Bh = np.random.default_rng(5100 + lg).uniform(-1, 1, n * n).astype(np.float32)
    blas.setGemmMode(mode)
    job = sharding.RowShardedGemm(blas, n, n, n, gather=False)                 # C stays sharded: no collective
    job.A.from_numpy(Ah)
    job.B.from_numpy(Bh)
    fused = None
    fused_err = None
    if world > 1:
        try:
            fused = sharding.RowShardedGemm(blas, n, n, n, gather=True, allow_multicast=not args.no_multicast,
                                            operands=(job.A, job.B))           # same operands, C in symmetric memory
        except Exception as e:                                                 # no peer access on this box
            fused_err = "%s: %s" % (type(e).__name__, e)
            log("fused gather unavailable:", fused_err)
    _capi.sync()

    def compute_step():
        blas.setGemmMode(mode)
        job.run()

    def fused_step():
        blas.setGemmMode(mode)
        fused.run()

    headline_step = fused_step if fused is not None else compute_step
    with sampler:
        _capi.launch_count(reset=True)
        nv0 = sampler.nvlink_bytes()
        total_ms = timed(headline_step, args.steps, args.warmup, tag="headline")
        nv1 = sampler.nvlink_bytes()
        launches = _capi.launch_count()
        gpu_launches = launches - launches * args.warmup // (args.steps + args.warmup)
        gemm_path = _capi.last_gemm_path()
        ms_per_step = total_ms / args.steps
        compute_only = None
        if fused is not None:
            c_ms = timed(compute_step, args.steps, args.warmup, tag="compute_only") / args.steps
            compute_only = {"value": 2.0 * n ** 3 / (c_ms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": c_ms,
                            "note": "same stripes, C left sharded: no data-path collective"}
    clocks = sampler.summary()
    flops_total = 2.0 * n * n * n
    flops_rank = flops_total / world
    value = flops_total / (ms_per_step * 1e-3) / 1e12
    tf32_meas = tf32_measured_peak()
    tf32_peak = tf32_meas["burst_tflops"] if tf32_meas else peaks["bf16_tflops"] / 2.0
    achieved = flops_rank / (ms_per_step * 1e-3) / 1e12
    achieved_note = ("algorithmic flops 2*N^3 (the kernel issues 3x that on the tensor pipe)" if args.mode == "tf32x3"
                     else "algorithmic flops 2*N^3")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": achieved / tf32_peak,
                # the ncu capture is of the single-GPU launch (the whole 16384^3 problem); stripes were not captured
                "traffic": ncu_traffic(gemm_path, n) if world == 1 else None, "kernel": gemm_path,
                "per": "GPU (this rank's stripe: 2*N^3/%d flop per launch)" % world,
                "peak_note": ("TF32 dense peak MEASURED on a pool B200 (profiles/peaks_tf32_r02.json: cuBLAS TF32 8192^3 "
                              "through torch, burst = best of 10); " if tf32_meas else
                              "TF32 dense = 1/2 of the measured cuBLAS bf16 burst figure (%s); " % peaks["source"])
                             + achieved_note,
                "frac_of_half_bf16_burst": achieved / (peaks["bf16_tflops"] / 2.0)}

    # the same launch back to back for ~2 s: the power-capped figure, against the sustained cuBLAS peak
    if not args.quick:
        reps = max(args.steps, int(2000.0 / max(ms_per_step, 1e-3)))
        sus_ms = timed(compute_step, reps, 0) / reps
        sus = flops_rank / (sus_ms * 1e-3) / 1e12
        sus_peak = tf32_meas["sustained_tflops"] if tf32_meas else peaks["bf16_tflops_sustained"] / 2.0
        roofline["sustained"] = {"achieved": sus, "peak": sus_peak, "frac": sus / sus_peak, "launches": reps,
                                 "ms_per_launch": sus_ms,
                                 "note": "compute-only launch repeated for ~2 s (sw_power_cap engaged); peak = the "
                                         "sustained cuBLAS TF32 figure measured the same way"}

    # ---- multi-GPU context: gemm + ncclAllGather of the stripes (library collective, not overlapped) ------------
    with_gather = None
    if world > 1:
        with_gather = {"compute_only": compute_only}
        if fused is not None:
            # every rank checks rows it did NOT compute against a float64 product of the owner's A (same seeds)
            peer = (rank + 1) % world
            Ap = np.random.default_rng(5000 + lg + 1000 * peer).uniform(-1, 1, rows * n).astype(np.float32).reshape(rows, n)
            _capi.sync()
            got = fused.C.read_range((peer * rows) * n, 4 * n).reshape(4, n)
            ref = Ap[:4].astype(np.float64) @ Bh.reshape(n, n).astype(np.float64)
            with_gather["fused"] = {
                "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "kernel": gemm_path,
                "multicast": bool(fused.multicast),
                "peer_rows_rel_err": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))),
                "nvlink_counters": (None if not (nv0 and nv1) else {
                    "tx_bytes_per_step": (nv1[0] - nv0[0]) / (args.steps + args.warmup),
                    "rx_bytes_per_step": (nv1[1] - nv0[1]) / (args.steps + args.warmup),
                    "tx_GBps_during_step": (nv1[0] - nv0[0]) / (args.steps + args.warmup) / (ms_per_step * 1e-3) / 1e9,
                    "rx_GBps_during_step": (nv1[1] - nv0[1]) / (args.steps + args.warmup) / (ms_per_step * 1e-3) / 1e9,
                    "note": "NVML NVLINK_THROUGHPUT_DATA_TX / RX of this rank's GPU (all links) read before and after "
                            "the headline region (warm-up steps included in the count)"}),
                "nvlink_bytes_out_per_rank": rows * n * 4 * (1 if fused.multicast else world - 1),
                "nvlink_bytes_in_per_rank": rows * n * 4 * (world - 1) + (rows * n * 4 if fused.multicast else 0),
                "note": "the headline: C stripe stored into every rank's full C from the GEMM epilogue (%s) + "
                        "symmetric-memory barrier" % ("NVLS multimem.st, one copy leaves the GPU" if fused.multicast
                                                       else "unicast NVLink peer stores")}
        else:
            with_gather["fused"] = {"unavailable": fused_err}
        try:
            tC = sharding.tensor_of(job.C, (rows, n))
            with torch.cuda.stream(stream):
                full = torch.empty((n, n), dtype=torch.float32, device="cuda")

            def nccl_step():
                compute_step()
                dist.all_gather_into_tensor(full, tC)

            g_steps = max(3, min(args.steps, 10))
            g_ms = timed(nccl_step, g_steps, 3, tag="nccl_gather") / g_steps
            with_gather["nccl"] = {"value": flops_total / (g_ms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": g_ms,
                                   "allgather_bytes_per_rank": (world - 1) * rows * n * 4,
                                   "note": "gemm + ncclAllGather of the row stripes of C, same stream (not overlapped); "
                                           "library collective, for context"}
            del full
        except Exception as e:  # pragma: no cover
            with_gather["nccl"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    # ---- e2e: the same job with HOST operands (sharding.RowShardedGemm.run_from_host) ---------------------------
    # each rank: its A stripe and its 1/world slice of B from pinned memory -> h2d; B slices all-gathered over NVLink;
    # transfer-pipelined multiply; C stripe d2h into pinned memory (the step's result), all inside the timed region
    hA, hB, hC = job.host_buffers()
    hA[...] = Ah.reshape(rows, n)
    hB[...] = Bh.reshape(n, n)[job.bstart:job.bstart + job.brows]
    pcie = measure_pcie(_capi, job) if (not args.quick or world > 1) else None
    if pcie and world > 1:
        # the same copies with ALL ranks going at once: what the host side (one socket's memory and I/O fabric serving
        # every GPU's link) sustains per rank when the links are used together, as the e2e step uses them
        barrier()
        conc = measure_pcie(_capi, job, barrier=barrier)
        pcie["concurrent_h2d_gbs"], pcie["concurrent_d2h_gbs"] = conc["h2d_gbs"], conc["d2h_gbs"]
    _capi.sync()

    def e2e_step():
        job.run_from_host()

    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = timed(e2e_step, e2e_steps, 3, tag="e2e") / e2e_steps
    h2d_b, d2h_b = int(hA.nbytes + hB.nbytes), int(hC.nbytes)
    e2e = {"value": flops_total / (e2e_ms * 1e-3) / 1e12, "unit": UNIT,
           "h2d_bytes_per_step": h2d_b * world, "d2h_bytes_per_step": d2h_b * world,
           "h2d_bytes_per_step_per_rank": h2d_b, "d2h_bytes_per_step_per_rank": d2h_b,
           "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "sharding.RowShardedGemm.run_from_host (service.blas().gemm with host-filled buffers on 1 GPU)",
           "note": "per rank: A stripe + 1/%d of B up, C stripe down; B's slices all-gathered over NVLink" % world}
    if pcie:
        bound_ms = max(h2d_b / (pcie["h2d_gbs"] * 1e6), d2h_b / (pcie["d2h_gbs"] * 1e6))
        e2e["pcie"] = pcie
        e2e["pcie_frac"] = bound_ms / e2e_ms
        if "concurrent_h2d_gbs" in pcie:
            cb = max(h2d_b / (pcie["concurrent_h2d_gbs"] * 1e6), d2h_b / (pcie["concurrent_d2h_gbs"] * 1e6))
            e2e["pcie_frac_concurrent"] = cb / e2e_ms
        e2e["pcie_note"] = ("max(h2d bytes / measured h2d GB/s, d2h bytes / measured d2h GB/s) / step time (this rank's link); "
                            "_concurrent: against the rates measured with every rank copying at once")
    got = hC[:2].astype(np.float64)
    ref = Ah.reshape(rows, n)[:2].astype(np.float64) @ Bh.reshape(n, n).astype(np.float64)
    e2e["rows_rel_err"] = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))

    # ---- other GEMM modes (co-headlines: the parity mode 3xTF32, FFMA, float64) + cuBLAS for context ------------
    gemm_modes = {}
    if not args.quick and world == 1:
        gemm_modes = run_gemm_modes(args, svc, timed, n, job, achieved, ms_per_step, stream)

    # ---- config C5: the sgemm sweep M=N=K in {1024 ... 16384} per GPU ------------------------------------------
    sweep = {}
    if not args.quick and not args.no_sweep and world == 1:
        sweep = run_sweep(args, svc, timed, n, gemm_modes, Ah, Bh)

    # ---- HBM-bound configs (C2, C3, C4): rotating buffer sets defeat L2 --------------------------------
    hbm_ops = {}
    if not args.quick:
        _capi.sync()
        time.sleep(2.0)              # the GEMM legs held the chip at its power cap: let the clock recover
        hbm_ops = run_hbm_ops(args, svc, timed, peaks, world, rank)

    # ---- cpu_baseline (rank 0, N=1 only): bounded sample of the same GEMM with the oracle ---------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        crows = 4
        t, _, Cref = reference_gemm_sample(n, crows, 1, oracle_inputs(Ah), oracle_inputs(Bh))
        compute_step()
        got = job.C.read_range(0, crows * n).reshape(crows, n).astype(np.float64)
        err = float(np.max(np.abs(got - Cref)) / np.max(np.abs(Cref)))
        cpu_baseline = {
            "value": 2.0 * crows * n * n / t / 1e12, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d of %d rows of C (N=K=%d), reference loop order m,n,k in double, 1 thread (the reference "
                      "reports 1 thread, PhpBlas.php:49-54); C restatement of PhpBlas::gemm -- the PHP interpreter "
                      "itself is not installable here and would be much slower" % (crows, n, n),
            "seconds": t, "gpu_vs_oracle_max_rel_err_on_sample": err,
        }
        try:                                            # numpy / OpenBLAS on all cores for context
            k = 4096
            a = Ah[: k * k].reshape(k, k)
            b = Bh[: k * k].reshape(k, k)
            a @ b
            t0 = time.perf_counter()
            a @ b
            t1 = time.perf_counter() - t0
            cpu_baseline["numpy_openblas_sgemm"] = {"value": 2.0 * k ** 3 / t1 / 1e12, "unit": UNIT,
                                                    "cores": os.cpu_count(), "n": k}
        except Exception:
            pass

    # ---- config C1: MatrixOperator::cross 256x256 float32 (the reference's own CPU-runnable case): latency -----
    cross_256 = run_cross_256(args, svc, stream, world) if (rank == 0 and not args.quick) else None

    # per-rank times and clocks (the max over ranks is what the headline uses)
    rank_clocks = [clocks]
    if world > 1:
        rank_clocks = [None] * world
        dist.all_gather_object(rank_clocks, clocks)
    log("rank %d/%d: %.3f ms/step, %.1f TFLOP/s aggregate" % (rank, world, ms_per_step, value))
    if rank == 0:
        co_headline = None
        if gemm_modes.get("tf32x3"):
            co_headline = {"metric": "sgemm TFLOP/s in the parity mode (3xTF32: inside the reference's isclose bound)",
                           "value": gemm_modes["tf32x3"]["tflops_per_gpu"], "unit": UNIT,
                           "ms_per_step": gemm_modes["tf32x3"]["ms"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"tf32": "tf32 (f32 in/out, f32 accumulate)", "tf32x3": "3xtf32 (f32 in/out, f32 accumulate)",
                      "fp32_simt": "f32"}[args.mode],
            "data": "synthetic", "config": workload_config(n, world, args.mode),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": gpu_launches,
            "clocks": clocks, "co_headline_parity_mode": co_headline,
            "per_rank": {"ms_per_step": per_rank, "clocks": rank_clocks, "numa_binding": numa},
            "gemm_modes": gemm_modes, "sgemm_sweep": sweep, "hbm_ops": hbm_ops,
            "with_gather": with_gather, "cross_256": cross_256, "peaks": peaks, "tf32_peak_measured": tf32_meas,
        }
        emit(line)
    if fused is not None:
        fused.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def tf32_measured_peak():
    """cuBLAS TF32 8192^3 burst / sustained measured on a pool B200 with tools/peak_probe.py (committed copy)."""
    try:
        with open(os.path.join(ROOT, "profiles", "peaks_tf32_r02.json")) as f:
            p = json.load(f)
        return {"burst_tflops": float(p["tf32"]["burst_tflops"]), "sustained_tflops": float(p["tf32"]["sustained_tflops"]),
                "fp32_tflops": float(p["fp32"]["burst_tflops"]), "fp64_tflops": float(p["fp64"]["burst_tflops"]),
                "source": "profiles/peaks_tf32_r02.json (tools/peak_probe.py: cuBLAS through torch, CUDA events)"}
    except Exception:
        return None


def measure_pcie(_capi, job, barrier=None):
    """This rank's PCIe link: one large pinned h2d, one d2h, timed with the host clock around synchronous copies
    (up to 256 MiB each, best of 3).  With `barrier` every rank starts each copy together."""
    import ctypes
    lib = _capi.load()
    dA, hA = job.A.raw_pointers()
    nbytes = min(job.A.bytes(), 256 << 20)
    out = {}
    for name, fn, args_ in (("h2d_gbs", lib.rb200_buffer_h2d, (dA, 0, hA, nbytes)),
                            ("d2h_gbs", lib.rb200_buffer_d2h, (hA, dA, 0, nbytes))):
        best = 1e30
        for _ in range(3):
            _capi.sync()
            if barrier is not None:
                barrier()
            t0 = time.perf_counter()
            _capi.check(fn(*args_))
            best = min(best, time.perf_counter() - t0)
        out[name] = nbytes / best / 1e9
    out["bytes"] = nbytes
    return out


def run_gemm_modes(args, svc, timed, n, job, achieved, ms_per_step, stream):
    """The other arithmetic modes on the headline operands (1 GPU): 3xTF32 (the parity mode), FFMA float32, DFMA
    float64 (north_star (a); its own random operands), plus cuBLAS TF32 / FP64 through torch for context."""
    import torch
    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import BLAS, CudaBuffer, dtypes
    blas = svc.blas()
    out = {}
    Cm = CudaBuffer(n * n, dtypes.float32)

    def mode_step(m):
        blas.setGemmMode(m)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, n, n, n, 1.0, job.A, 0, n, job.B, 0, n, 0.0, Cm, 0, n)

    for name, mm in (("tf32", rb.GEMM_TF32), ("tf32x3", rb.GEMM_TF32X3), ("fp32_simt", rb.GEMM_FP32_SIMT)):
        if name == args.mode:
            out[name] = {"tflops_per_gpu": achieved, "ms": ms_per_step}
            continue
        k = 2 if name == "fp32_simt" else max(3, min(args.steps, 10))
        ms = timed(lambda mm=mm: mode_step(mm), k, 1 if name == "fp32_simt" else 2) / k
        out[name] = {"tflops_per_gpu": 2.0 * n ** 3 / (ms * 1e-3) / 1e12, "ms": ms}
    peak = tf32_measured_peak()
    if peak:
        out["fp32_simt"]["frac_of_cublas_fp32"] = out["fp32_simt"]["tflops_per_gpu"] / peak["fp32_tflops"]
    del Cm
    # float64 (rb200_dgemm) at 8192^3
    nd = 8192
    rng = np.random.default_rng(77)
    Ad = CudaBuffer(nd * nd, dtypes.float64, zero=False).from_numpy(rng.uniform(-1, 1, nd * nd))
    Bd = CudaBuffer(nd * nd, dtypes.float64, zero=False).from_numpy(rng.uniform(-1, 1, nd * nd))
    Cd = CudaBuffer(nd * nd, dtypes.float64)

    def f64_step():
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, nd, nd, nd, 1.0, Ad, 0, nd, Bd, 0, nd, 0.0, Cd, 0, nd)
    ms = timed(f64_step, 2, 1) / 2
    from rindow_math_matrix_b200 import _capi
    out["f64"] = {"tflops_per_gpu": 2.0 * nd ** 3 / (ms * 1e-3) / 1e12, "ms": ms, "n": nd, "path": _capi.last_gemm_path()}
    if peak:
        out["f64"]["frac_of_cublas_fp64"] = out["f64"]["tflops_per_gpu"] / peak["fp64_tflops"]
    del Ad, Bd, Cd
    try:
        ta = torch.as_tensor(job.A, device="cuda").view(n, n)
        tb = torch.as_tensor(job.B, device="cuda").view(n, n)
        torch.backends.cuda.matmul.allow_tf32 = True
        with torch.cuda.stream(stream):
            tc = torch.empty(n, n, device="cuda")
        k = max(3, min(args.steps, 10))
        ms = timed(lambda: torch.matmul(ta, tb, out=tc), k, 2) / k
        out["cublas_tf32_via_torch"] = {"tflops_per_gpu": 2.0 * n ** 3 / (ms * 1e-3) / 1e12, "ms": ms,
                                        "note": "library GEMM for context, not on the product path"}
    except Exception as e:  # pragma: no cover
        log("cuBLAS TF32 context measurement failed:", e)
    blas.setGemmMode({"tf32": rb.GEMM_TF32, "tf32x3": rb.GEMM_TF32X3, "fp32_simt": rb.GEMM_FP32_SIMT}[args.mode])
    return out


def run_sweep(args, svc, timed, n, gemm_modes, Ah, Bh):
    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import BLAS, CudaBuffer, dtypes
    blas = svc.blas()
    sweep = {}
    from rindow_math_matrix_b200 import _capi as _c
    _c.sync()
    time.sleep(2.0)        # the FFMA / DFMA legs before this one held the chip at its power cap: let the clock recover
    for sz in (1024, 2048, 4096, 8192, 16384):
        if sz == n:
            sweep[str(sz)] = {"tf32_tflops": gemm_modes.get("tf32", {}).get("tflops_per_gpu"),
                              "tf32x3_tflops": gemm_modes.get("tf32x3", {}).get("tflops_per_gpu")}
            continue
        sa = CudaBuffer(sz * sz, dtypes.float32, zero=False)
        sb = CudaBuffer(sz * sz, dtypes.float32, zero=False)
        sc = CudaBuffer(sz * sz, dtypes.float32, zero=False)
        # uniform(-1,1) operands like the headline run: constant operands toggle fewer bits, draw less
        # power and clock ~35 % higher on this part -- such a number would not be comparable
        sa.from_numpy(Ah[:sz * sz])
        sb.from_numpy(Bh[:sz * sz])
        row = {}
        from rindow_math_matrix_b200 import _capi
        for name, mm in (("tf32", rb.GEMM_TF32), ("tf32x3", rb.GEMM_TF32X3)):
            blas.setGemmMode(mm)
            k = 20 if sz <= 4096 else 5

            def sweep_step():
                blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, sz, sz, sz, 1.0, sa, 0, sz, sb, 0, sz, 0.0,
                          sc, 0, sz)
            ms = timed(sweep_step, k, 3) / k
            row[name + "_issued_tflops"] = 2.0 * sz ** 3 / (ms * 1e-3) / 1e12
            row[name + "_issued_ms"] = ms
            if sz <= 4096:
                # launches of a few microseconds: the host cost of issuing each call (Python + ctypes + tensor-map
                # encoding, ~20 us) hides the kernel.  The same k calls recorded once in a CUDA graph (rb200_graph_*)
                # and replayed measure what the GPU does; that figure is the sweep's number for these sizes
                g = _capi.Graph()
                with g:
                    for _ in range(k):
                        sweep_step()
                gms = timed(g.launch, 5, 3) / (5 * k)
                g.close()
                row[name + "_tflops"] = 2.0 * sz ** 3 / (gms * 1e-3) / 1e12
                row[name + "_ms"] = gms
                row["timing"] = "CUDA-graph replay of %d calls (issued call by call: *_issued_*)" % k
            else:
                row[name + "_tflops"] = row[name + "_issued_tflops"]
                row[name + "_ms"] = ms
        from rindow_math_matrix_b200 import _capi
        row["path"] = _capi.last_gemm_path()
        sweep[str(sz)] = row
        del sa, sb, sc
    return sweep


def run_cross_256(args, svc, stream, world):
    import torch
    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import _capi
    try:
        from rindow_math_matrix_b200 import MatrixOperator
        mo = MatrixOperator(svc)
        svc.blas().setGemmMode(rb.GEMM_AUTO)
        a256 = np.random.default_rng(1234).uniform(-1, 1, (256, 256)).astype(np.float32)
        b256 = np.random.default_rng(1235).uniform(-1, 1, (256, 256)).astype(np.float32)
        A2, B2 = mo.array(a256), mo.array(b256)
        reps = 50
        with torch.cuda.stream(stream):
            for _ in range(5):
                mo.cross(A2, B2)
            _capi.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                out256 = mo.cross(A2, B2)
            _capi.sync()
            host_us = (time.perf_counter() - t0) / reps * 1e6
        got = out256.to_numpy().astype(np.float64)
        ref = a256.astype(np.float64) @ b256.astype(np.float64)
        cross_256 = {"us_per_call": host_us, "calls": reps, "path": _capi.last_gemm_path(),
                     "max_rel_err": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))),
                     "note": "MatrixOperator.cross on device-resident operands, wall clock per call incl. the "
                             "zero-filled C allocation the reference does (MatrixOperator.php:479-503); launch bound"}
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            co = np.zeros(256 * 256)
            t0 = time.perf_counter()
            oracle.gemm_rows(0, 256, 256, 256, 1.0, oracle.as_buffer(a256), 256, oracle.as_buffer(b256), 256, 1.0, co, 256)
            cross_256["oracle_port_us"] = (time.perf_counter() - t0) * 1e6
        return cross_256
    except Exception as e:  # pragma: no cover
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def oracle_inputs(a32):
    import oracle
    return oracle.as_buffer(a32)


def run_hbm_ops(args, svc, timed, peaks, world, rank):
    """C2/C3/C4 at BASELINE sizes; each op cycles through 3 operand sets (> L2) so every launch
    streams from HBM.  GB/s uses the ALGORITHMIC bytes of SURVEY.md section 8(d)."""
    from rindow_math_matrix_b200 import CudaBuffer, dtypes
    math = svc.math()
    out = {}
    sets = 3
    steps = max(6, min(args.steps, 30))
    steps -= steps % sets
    hbm = peaks["hbm_gbs"]

    def report(name, fn, bytes_per_call, extra=None):
        state = {"i": 0}

        def step():
            fn(state["i"] % sets)
            state["i"] += 1
        ms = timed(step, steps, sets) / steps
        gbs = bytes_per_call / (ms * 1e-3) / 1e9
        out[name] = {"GB/s_per_gpu": gbs, "GB/s": gbs * world, "frac_of_measured_hbm": gbs / hbm,
                     "frac_of_nominal_8TBs": gbs / 8000.0, "ms": ms, "algorithmic_bytes": bytes_per_call}
        if extra:
            out[name].update(extra)

    rng = np.random.default_rng(2001 + rank)
    # C2: add / multiply f32 [8192,4096] (+|x) [4096]
    m, n = 8192, 4096
    a0 = rng.uniform(-1, 1, m * n).astype(np.float32)
    As = [CudaBuffer(m * n, dtypes.float32, zero=False).from_numpy(a0) for _ in range(sets)]
    X = CudaBuffer(n, dtypes.float32, zero=False).from_numpy(rng.uniform(0.9, 1.1, n).astype(np.float32))
    ew_bytes = 2 * m * n * 4 + n * 4
    report("add", lambda i: math.add(False, m, n, 1.0, X, 0, 1, As[i], 0, n), ew_bytes,
           {"shape": "[8192,4096] + [4096]"})
    report("multiply", lambda i: math.multiply(False, m, n, X, 0, 1, As[i], 0, n), ew_bytes,
           {"shape": "[8192,4096] * [4096]"})
    # section 8(f) widening on the same [8192,4096] operands: axpy / copy / gemv / transpose (omatcopy)
    from rindow_math_matrix_b200 import BLAS as _B
    blas1 = svc.blas()
    Y2 = CudaBuffer(m * n, dtypes.float32, zero=False).from_numpy(a0)
    report("axpy", lambda i: blas1.axpy(m * n, 0.5, As[i], 0, 1, Y2, 0, 1), 3 * m * n * 4, {"shape": "n=33554432"})
    report("copy", lambda i: blas1.copy(m * n, As[i], 0, 1, Y2, 0, 1), 2 * m * n * 4, {"shape": "n=33554432"})
    report("transpose", lambda i: blas1.omatcopy(_B.RowMajor, _B.Trans, m, n, 1.0, As[i], 0, n, Y2, 0, m),
           2 * m * n * 4, {"shape": "[8192,4096] -> [4096,8192] (omatcopy)"})
    report("slice", lambda i: math.slice(False, False, m, n, 1, 1, As[i], 0, 1, Y2, 0, 1, 0, m, 512, 3072, 0, 1),
           2 * m * 3072 * 4, {"shape": "[8192,4096] -> columns 512..3583 (split / concat along axis 1)"})
    report("repeat", lambda i: math.repeat(m, 1024, 4, As[i], 0, Y2, 0), (m * 1024 + m * 4096) * 4,
           {"shape": "[8192,1024] -> [8192,4,1024]"})
    # attention-score padding mask over [64 batch, 16 heads x 128 queries, 256 keys] (the same 128 MB operands): per batch
    # the keys past a random length are set to -1e9; only those elements are written
    lens = np.random.default_rng(77).integers(64, 257, 64)
    mk = (np.arange(256)[None, :] < lens[:, None]).astype(np.uint8)
    Mk = CudaBuffer(64 * 256, dtypes.bool_).from_numpy(mk.reshape(-1))
    masked = int((mk == 0).sum()) * 16 * 128
    report("masking", lambda i: math.masking(64, 16 * 128, 256, 1, -1e9, 0, Mk, 0, As[i], 0), masked * 4,
           {"shape": "scores [64,2048,256], key padding mask [64,256], %.0f %% masked (bytes = masked elements written)"
                     % (100.0 * masked / (m * n))})
    report("cumsum_rows", lambda i: math.cumsumb(m, n, 1, As[i], 0, False, False, Y2, 0), 2 * m * n * 4,
           {"shape": "cumsum([8192,4096], axis=1): one pass"})
    report("cumsum_cols", lambda i: math.cumsumb(1, m, n, As[i], 0, False, False, Y2, 0), 2 * m * n * 4,
           {"shape": "cumsum([8192,4096], axis=0): totals pass + scan pass, 2 reads + 1 write for 1 + 1 algorithmic"})
    Yv = CudaBuffer(m, dtypes.float32)
    report("gemv", lambda i: blas1.gemv(_B.RowMajor, _B.NoTrans, m, n, 1.0, As[i], 0, n, X, 0, 1, 0.0, Yv, 0, 1),
           m * n * 4 + n * 4 + m * 4, {"shape": "[8192,4096] x [4096]"})
    Yt = CudaBuffer(n, dtypes.float32)
    Xt = CudaBuffer(m, dtypes.float32, zero=False).from_numpy(rng.uniform(-1, 1, m).astype(np.float32))
    report("gemv_trans", lambda i: blas1.gemv(_B.RowMajor, _B.Trans, m, n, 1.0, As[i], 0, n, Xt, 0, 1, 0.0, Yt, 0, 1),
           m * n * 4 + n * 4 + m * 4, {"shape": "[8192,4096]^T x [8192]"})
    # section 8(f) ranks 2 and 4 on the same operands: broadcast / unary families, astype, N-D transpose, gather
    report("maximum", lambda i: math.maximum(m, n, As[i], 0, n, X, 0, 1), ew_bytes, {"shape": "max([8192,4096], [4096])"})
    report("greater", lambda i: math.greater(m, n, As[i], 0, n, X, 0, 1), ew_bytes, {"shape": "[8192,4096] > [4096]"})
    report("exp", lambda i: math.exp(m * n, As[i], 0, 1), 2 * m * n * 4, {"shape": "n=33554432 in place"})
    report("increment", lambda i: math.increment(m * n, 1.0001, As[i], 0, 1, 0.001), 2 * m * n * 4,
           {"shape": "n=33554432 in place"})
    I32 = CudaBuffer(m * n, dtypes.int32, zero=False)
    report("astype_f32_i32", lambda i: math.astype(m * n, dtypes.int32, As[i], 0, 1, I32, 0, 1), 2 * m * n * 4,
           {"shape": "n=33554432 float32 -> int32"})
    del I32
    sh3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([64, 512, 1024], np.int32))
    pm3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([0, 2, 1], np.int32))
    report("transpose_nd", lambda i: math.transpose(sh3, pm3, As[i], 0, Y2, 0), 2 * m * n * 4,
           {"shape": "[64,512,1024] -> [64,1024,512] (perm 0,2,1)"})
    vocab, dim, ntok = 32768, 1024, 32768
    ids = CudaBuffer(ntok, dtypes.int32).from_numpy(rng.integers(0, vocab, ntok).astype(np.int32))
    report("gather", lambda i: math.gather(False, False, ntok, dim, vocab, ids, 0, As[i], 0, Y2, 0), 2 * ntok * dim * 4,
           {"shape": "table [32768,1024], 32768 row indices (embedding lookup)"})
    report("scatter_add", lambda i: math.gather(True, True, ntok, dim, vocab, ids, 0, As[i], 0, Y2, 0), 3 * ntok * dim * 4,
           {"shape": "table [32768,1024] += 32768 rows (embedding gradient; sort + in-order sums, 4 launches)"})
    del As, Y2
    # C3: reductions / softmax on f32 [65536,1024] axis=1
    m, n = 65536, 1024
    r0 = np.random.default_rng(3001 + rank).uniform(-1, 1, m * n).astype(np.float32)
    Rs = [CudaBuffer(m * n, dtypes.float32, zero=False).from_numpy(r0) for _ in range(sets)]
    Bf = CudaBuffer(m, dtypes.float32)
    Bi = CudaBuffer(m, dtypes.int32)
    rd_bytes = m * n * 4 + m * 4
    report("reduceSum", lambda i: math.reduceSum(m, n, 1, Rs[i], 0, Bf, 0), rd_bytes, {"shape": "[65536,1024] axis=1"})
    report("reduceMax", lambda i: math.reduceMax(m, n, 1, Rs[i], 0, Bf, 0), rd_bytes, {"shape": "[65536,1024] axis=1"})
    report("reduceArgMax", lambda i: math.reduceArgMax(m, n, 1, Rs[i], 0, Bi, 0), rd_bytes,
           {"shape": "[65536,1024] axis=1"})
    report("reduceSum_axis0", lambda i: math.reduceSum(1, m, n, Rs[i], 0, Bf, 0), m * n * 4 + n * 4,
           {"shape": "[65536,1024] axis=0 (split + merge path)"})
    report("softmax", lambda i: math.softmax(m, n, Rs[i], 0, n), 2 * m * n * 4, {"shape": "[65536,1024] in place"})
    del Rs
    # Conv3D sibling of C4: clips [8,16,112,112,3], 3x3x3 valid -> cols [8,14,110,110,3,3,3,3] (440 MB)
    b3, d3, h3, w3, c3 = 8, 16, 112, 112, 3
    vol = np.random.default_rng(4101 + rank).uniform(0, 1, b3 * d3 * h3 * w3 * c3).astype(np.float32)
    V3 = [CudaBuffer(vol.size, dtypes.float32, zero=False).from_numpy(vol) for _ in range(sets)]
    n3 = b3 * (d3 - 2) * (h3 - 2) * (w3 - 2) * 81
    C3 = CudaBuffer(n3, dtypes.float32, zero=False)
    report("im2col3d", lambda i: math.im2col3d(False, V3[i], 0, vol.size, b3, d3, h3, w3, c3, 3, 3, 3, 1, 1, 1, False, False,
                                                1, 1, 1, False, C3, 0, n3), (vol.size + n3) * 4,
           {"shape": "8x16x112x112x3, 3x3x3 valid -> [8,14,110,110,3,3,3,3]"})
    del V3, C3
    # C4: im2col2d 64x224x224x3, 3x3, stride 1, valid padding
    b, h, w, c = 64, 224, 224, 3
    img = np.random.default_rng(4001 + rank).uniform(0, 1, b * h * w * c).astype(np.float32)
    images = CudaBuffer(img.size, dtypes.float32, zero=False).from_numpy(img)
    ncols = b * 222 * 222 * 27
    cols = [CudaBuffer(ncols, dtypes.float32, zero=False) for _ in range(2)]
    ic_bytes = img.size * 4 + ncols * 4
    sets_saved = sets
    sets = 2
    steps -= steps % sets
    report("im2col2d", lambda i: math.im2col2d(False, images, 0, img.size, b, h, w, c, 3, 3, 1, 1, False, False, 1, 1,
                                                False, cols[i], 0, ncols), ic_bytes,
           {"shape": "64x224x224x3, 3x3 valid -> [64,222,222,3,3,3]"})
    # conv gemm [3154176 x 27] x [27 x 64] (HBM-bound: report GB/s on the per-call bytes)
    from rindow_math_matrix_b200 import BLAS
    blas = svc.blas()
    M = b * 222 * 222
    Wt = CudaBuffer(27 * 64, dtypes.float32, zero=False).from_numpy(
        (np.random.default_rng(4002).standard_normal(27 * 64) * 0.1).astype(np.float32))
    outs = [CudaBuffer(M * 64, dtypes.float32, zero=False) for _ in range(2)]
    from rindow_math_matrix_b200 import _capi
    blas.setGemmMode(0)
    conv_bytes = ncols * 4 + 27 * 64 * 4 + M * 64 * 4
    report("conv_gemm", lambda i: blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, 64, 27, 1.0, cols[i], 0, 27,
                                            Wt, 0, 64, 0.0, outs[i], 0, 64), conv_bytes,
           {"shape": "[3154176,27] x [27,64]", "flops": 2.0 * M * 27 * 64})
    out["conv_gemm"]["path"] = _capi.last_gemm_path()
    # fused conv2d forward (implicit GEMM): images + W + out only -- 846,011,136 B
    fused_bytes = img.size * 4 + 27 * 64 * 4 + M * 64 * 4
    ok = {"v": True}

    def fused(i):
        ok["v"] = math.conv2d_forward(images, 0, b, h, w, c, 3, 3, 1, 1, False, 1, 1, Wt, 0, 64, None, 0,
                                      outs[i], 0) and ok["v"]
    report("conv2d_fused", fused, fused_bytes,
           {"shape": "64x224x224x3 * 3x3x3x64 -> [64,222,222,64] in one kernel (no cols)",
            "flops": 2.0 * M * 27 * 64, "two_call_bytes": ic_bytes + conv_bytes})
    out["conv2d_fused"]["path"] = _capi.last_gemm_path()
    out["conv2d_fused"]["fused_kernel_used"] = bool(ok["v"])
    sets = sets_saved
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            hbm_cpu_baselines(out)
        except Exception as e:  # pragma: no cover
            log("hbm cpu baselines failed:", e)
        # e2e of one HBM op through HOST buffers: A lives in its pinned mirror, add runs on the device, A is read back
        try:
            m, n = 8192, 4096
            Ae = CudaBuffer(m * n, dtypes.float32, zero=False)
            Ae.host()[:] = 0.5
            Xe = CudaBuffer(n, dtypes.float32, zero=False).from_numpy(np.full(n, 0.25, np.float32))

            def e2e_add(_i):
                Ae.mark_host_written()
                math.add(False, m, n, 1.0, Xe, 0, 1, Ae, 0, n)
                Ae.host()
            st = {"i": 0}
            ms = timed(lambda: e2e_add(0), 6, 3) / 6
            out["add"]["e2e"] = {"GB/s": (2 * m * n * 4 + n * 4) / (ms * 1e-3) / 1e9, "ms": ms,
                                 "h2d_bytes_per_step": m * n * 4, "d2h_bytes_per_step": m * n * 4,
                                 "note": "math().add with A current only in its pinned host mirror and read back from the "
                                         "host afterwards (h2d + kernel + d2h inside the timed region): PCIe-bound"}
        except Exception as e:  # pragma: no cover
            log("add e2e failed:", e)
    return out


def hbm_cpu_baselines(out):
    """B0 (the oracle: C restatement of the PhpMath loops, double, 1 thread, bounded rows of the same workload) and B1
    (numpy on the host, vectorised, full size) beside the HBM-bound ops of configs C2/C3/C4.
    GB/s on the ALGORITHMIC float32 bytes of the sample, like the GPU rows, so the ratios are like for like."""
    import oracle

    def t_of(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best

    rng = np.random.default_rng(2001)
    label = "C restatement of the reference PHP loop (double, 1 thread; the PHP interpreter is unavailable and slower)"
    # C2 add / multiply: 1024 of 8192 rows
    m, n = 1024, 4096
    a = oracle.as_buffer(rng.uniform(-1, 1, m * n).astype(np.float32))
    x = oracle.as_buffer(rng.uniform(0.9, 1.1, n).astype(np.float32))
    full32 = rng.uniform(-1, 1, 8192 * 4096).astype(np.float32).reshape(8192, 4096)
    x32 = x.astype(np.float32)
    byt = 2 * m * n * 4 + n * 4
    for name, fn, npfn in (("add", lambda: oracle.add(False, m, n, 1.0, x, 0, 1, a, 0, n), lambda: np.add(full32, x32, out=full32)),
                           ("multiply", lambda: oracle.multiply(False, m, n, x, 0, 1, a, 0, n), lambda: np.multiply(full32, x32, out=full32))):
        t = t_of(fn)
        tn = t_of(npfn)
        out[name]["cpu_baseline"] = {"value": byt / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                     "sample": "%d of 8192 rows of [8192,4096]; %s" % (m, label),
                                     "numpy": {"value": (2 * 8192 * 4096 * 4 + n * 4) / tn / 1e9, "unit": "GB/s", "cores": 1,
                                               "sample": "full size, in place, numpy vectorised (one thread)"}}
    del full32
    # C3 reductions / softmax: 4096 of 65536 rows
    m, n = 4096, 1024
    r32 = rng.uniform(-1, 1, m * n).astype(np.float32)
    r = oracle.as_buffer(r32)
    bsum, bidx = np.zeros(m), np.zeros(m, np.int64)
    r2 = r32.reshape(m, n)
    for name, fn, npfn, byt in (
            ("reduceSum", lambda: oracle.reduce_sum(m, n, 1, r, 0, bsum, 0), lambda: r2.sum(axis=1), m * n * 4 + m * 4),
            ("reduceMax", lambda: oracle.reduce_max(m, n, 1, r, 0, bsum, 0, sticky=True), lambda: r2.max(axis=1), m * n * 4 + m * 4),
            ("reduceArgMax", lambda: oracle.reduce_argmax(m, n, 1, r, 0, bidx, 0), lambda: r2.argmax(axis=1), m * n * 4 + m * 4),
            ("softmax", lambda: oracle.softmax(m, n, r.copy(), 0, n),
             lambda: (lambda e: e / e.sum(axis=1, keepdims=True))(np.exp(r2 - r2.max(axis=1, keepdims=True))), 2 * m * n * 4)):
        t = t_of(fn)
        tn = t_of(npfn)
        out[name]["cpu_baseline"] = {"value": byt / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                     "sample": "%d of 65536 rows of [65536,1024]; %s" % (m, label),
                                     "numpy": {"value": byt / tn / 1e9, "unit": "GB/s", "cores": 1,
                                               "sample": "the same %d rows, numpy vectorised (one thread)" % m}}
    # C4 im2col2d: 2 of 64 images
    b, h, w, c = 2, 224, 224, 3
    img32 = rng.uniform(0, 1, b * h * w * c).astype(np.float32)
    img = oracle.as_buffer(img32)
    ncols = b * 222 * 222 * 27
    cols = np.zeros(ncols)
    byt = (img.size + ncols) * 4
    t = t_of(lambda: oracle.im2col2d(False, img, 0, img.size, b, h, w, c, 3, 3, 1, 1, False, False, 1, 1, False, cols, 0, ncols))
    i4 = img32.reshape(b, h, w, c)
    tn = t_of(lambda: np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(i4, (3, 3), axis=(1, 2)).transpose(0, 1, 2, 4, 5, 3)))
    out["im2col2d"]["cpu_baseline"] = {"value": byt / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                       "sample": "%d of 64 images; %s" % (b, label),
                                       "numpy": {"value": byt / tn / 1e9, "unit": "GB/s", "cores": 1,
                                                 "sample": "the same images, stride-trick im2col + copy (one thread)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=16384, help="global sgemm size (M=N=K), row-sharded over the GPUs")
    ap.add_argument("--no-multicast", action="store_true", help="fused gather through unicast peer stores even when NVLS multicast is available")
    ap.add_argument("--mode", default="tf32", choices=["tf32", "tf32x3", "fp32_simt"])
    ap.add_argument("--quick", action="store_true", help="headline workload only (no hbm_ops / other modes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the C5 size sweep")
    args = ap.parse_args()
    if args.warmup < 3:
        log("bench.py: raising --warmup to the required minimum of 3")
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
