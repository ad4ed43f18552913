"""Host <-> device copy rates of this box for the sizes the C ABI moves (one N-vector = 1.6 MB, a probe batch = 48 MB), from pageable
and from pinned host memory (cudaMemcpyAsync + synchronize, median of 20)."""
import time
import torch
for nbytes in (1_600_000, 4_800_000, 48_000_000):
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for kind in ("pageable", "pinned"):
        hbuf = torch.empty(nbytes, dtype=torch.uint8)
        if kind == "pinned":
            hbuf = hbuf.pin_memory()
        hbuf.fill_(1)
        for direction in ("h2d", "d2h"):
            ts = []
            for _ in range(20):
                torch.cuda.synchronize(); t = time.perf_counter()
                if direction == "h2d":
                    d.copy_(hbuf, non_blocking=True)
                else:
                    hbuf.copy_(d, non_blocking=True)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
            ts.sort()
            print("%9d bytes %-8s %s: %.3f ms = %.1f GB/s" % (nbytes, kind, direction, 1e3 * ts[10], nbytes / ts[10] / 1e9))
