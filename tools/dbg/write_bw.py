"""Pure-write HBM bandwidth with incompressible data: broadcast an L2-resident random block into a large buffer."""
import torch
dev = "cuda"
small = torch.randn(2 * 1024 * 1024, device=dev)            # 8 MB, stays in L2
for total_mb in (807, 1024, 2048):
    reps = total_mb * 1024 * 1024 // (small.numel() * 4)
    out = torch.empty(reps, small.numel(), device=dev)
    src = small.expand(reps, small.numel())
    for _ in range(3):
        out.copy_(src)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    nbytes = out.numel() * 4
    print("write-only %5d MB: %.4f ms  %.0f GB/s" % (total_mb, best, nbytes / best / 1e6), flush=True)
    z = torch.zeros_like(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out.zero_()
    e1.record()
    torch.cuda.synchronize()
    print("zero_      %5d MB: %.4f ms  %.0f GB/s" % (total_mb, e0.elapsed_time(e1) / 10, nbytes / (e0.elapsed_time(e1) / 10) / 1e6), flush=True)
    del out, z
