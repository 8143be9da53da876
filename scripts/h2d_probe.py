"""Does splitting one large pinned host -> device copy over several streams raise the PCIe rate?  (e2e of the drop-in
call is 80 GB of upload.)"""
import time
import torch

n = 1 << 30  # 8 GiB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c = n // parts
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("H2D 8 GiB pinned, %d stream(s): %.3f s = %.1f GB/s" % (parts, best, n * 8 / best * 1e-9), flush=True)
