"""Pinned D2H / H2D bandwidth per process (run under torchrun to see contention between GPUs)."""
import os, time, torch
r = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(r)
n = 87 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 20
    print(f"rank {r} {name}: {n / t / 1e9:.1f} GB/s ({t*1e3:.2f} ms per 87 MB)", flush=True)
