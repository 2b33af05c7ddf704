"""HBM ceilings on this box for the access pattern of the env step (pure streaming write)."""
import torch
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for mb in (347, 1024, 2776):
    nbytes = mb * 1000 * 1000 // 16 * 16
    a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    t = timeit(lambda: a.zero_())
    print("fill  %5d MB: %.1f us  %.0f GB/s (write only)" % (mb, t * 1e3, nbytes / t / 1e6))
    t = timeit(lambda: a.view(torch.int32).fill_(7))
    print("fill32 %4d MB: %.1f us  %.0f GB/s (write only)" % (mb, t * 1e3, nbytes / t / 1e6))
    t = timeit(lambda: b.copy_(a))
    print("copy  %5d MB: %.1f us  %.0f GB/s (read+write)" % (mb, t * 1e3, 2 * nbytes / t / 1e6))
    t = timeit(lambda: torch.cuda.memset if False else a.view(torch.int64).sum())
    print("read  %5d MB: %.1f us  %.0f GB/s (read only)" % (mb, t * 1e3, nbytes / t / 1e6))

# Sustained write-only rate: the same buffer filled back to back with no synchronisation in between, the way the env
# step overwrites its observation buffer every launch (a single fill ends with up to ~126 MB still dirty in the L2, so
# its time understates the DRAM work; in steady state every byte of a fill has to be written back).
for mb in (347, 2776):
    nbytes = mb * 1000 * 1000 // 16 * 16
    a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    v = a.view(torch.int32)
    for _ in range(5): v.fill_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if mb < 1000 else 40
    e0.record()
    for i in range(reps): v.fill_(i)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    print("fill32 %4d MB back to back x%d: %.1f us per fill  %.0f GB/s (write only, sustained)" % (mb, reps, t * 1e3, nbytes / t / 1e6))
