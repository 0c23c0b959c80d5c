"""D2H / H2D rate of this box for a few transfer sizes (pinned host memory)."""
import time, torch
for mb in (0.8, 1.6, 3.2, 6.4, 12.8, 32, 160):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        t = time.perf_counter()
        reps = 20
        for _ in range(reps):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / reps
        print("%s %6.1f MB: %7.1f us  %5.1f GB/s" % (direction, mb, dt * 1e6, n / dt / 1e9))
