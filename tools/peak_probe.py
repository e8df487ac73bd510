#!/usr/bin/env python3
"""On-box denominators the roofline fractions of the GEMM family are quoted against (SURVEY.md section 8(d) asks for a
TF32 peak MEASURED on the box; MEASURED_PEAKS.json only has bf16): cuBLAS through torch.matmul at 8192^3 --
TF32 (allow_tf32), FP32 without tensor cores, FP64 -- best of 10 launches (burst) and back to back for 4 s
(sustained, the power-capped figure), timed with CUDA events, random U(-1,1) operands.  Library kernels, context
only: nothing here is on the product path.  Writes one JSON object to stdout (and to argv[1] if given)."""
import json
import sys
import time

import torch


def measure(dtype, n, allow_tf32, seconds=4.0):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.empty(n, n, device="cuda", dtype=dtype).uniform_(-1, 1, generator=g)
    b = torch.empty(n, n, device="cuda", dtype=dtype).uniform_(-1, 1, generator=g)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        time.sleep(0.05)
    reps = max(5, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / reps
    return {"n": n, "burst_tflops": flops / best / 1e9, "sustained_tflops": flops / sus / 1e9, "burst_ms": best,
            "sustained_ms": sus, "sustained_launches": reps}


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "how": __doc__.split("\n\n")[0]}
    out["tf32"] = measure(torch.float32, 8192, True)
    time.sleep(2)
    out["tf32_16384"] = measure(torch.float32, 16384, True, seconds=3.0)
    time.sleep(2)
    out["fp32"] = measure(torch.float32, 8192, False, seconds=3.0)
    time.sleep(2)
    out["fp64"] = measure(torch.float64, 8192, False, seconds=3.0)
    s = json.dumps(out, indent=1)
    print(s)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
