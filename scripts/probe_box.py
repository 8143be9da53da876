"""One-off hardware probe of the GPU box (host RAM/cores, fp64 GEMM peak, copy bandwidth)."""
import json, os, subprocess, time
import torch
out = {}
out["nproc"] = os.cpu_count()
out["meminfo"] = open("/proc/meminfo").read().split("\n")[:3]
out["lscpu"] = subprocess.run("lscpu | head -20", shell=True, capture_output=True, text=True).stdout
out["nvidia_smi"] = subprocess.run("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv", shell=True, capture_output=True, text=True).stdout
out["topo"] = subprocess.run("nvidia-smi topo -m | head -20", shell=True, capture_output=True, text=True).stdout
dev = torch.device("cuda:0")
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best
# fp64 GEMM peak (cuBLAS DGEMM) -- denominator only
for (m, n, k) in [(8192, 8192, 8192), (16384, 64, 16384), (16384, 32, 16384), (32768, 128, 32768)]:
    a = torch.randn(m, k, dtype=torch.float64, device=dev); b = torch.randn(k, n, dtype=torch.float64, device=dev)
    t = timeit(lambda: torch.matmul(a, b))
    out[f"dgemm_{m}x{n}x{k}_tflops"] = 2.0 * m * n * k / t / 1e12
    out[f"dgemm_{m}x{n}x{k}_GBps"] = 8.0 * (m * k + k * n + m * n) / t / 1e9
    del a, b
# copy bandwidth
x = torch.empty(1 << 30, dtype=torch.float64, device=dev); y = torch.empty_like(x)
t = timeit(lambda: y.copy_(x)); out["copy_GBps"] = 2 * x.numel() * 8 / t / 1e9
t = timeit(lambda: x.sum()); out["read_sum_GBps"] = x.numel() * 8 / t / 1e9
del x, y
# pinned H2D bandwidth
h = torch.empty(1 << 28, dtype=torch.float64).pin_memory(); d = torch.empty(1 << 28, dtype=torch.float64, device=dev)
t = timeit(lambda: d.copy_(h, non_blocking=True)); out["h2d_pinned_GBps"] = h.numel() * 8 / t / 1e9
t = timeit(lambda: h.copy_(d, non_blocking=True)); out["d2h_pinned_GBps"] = h.numel() * 8 / t / 1e9
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_box.json", "w"), indent=1)
