"""Write-only and read-only HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (dev helper)."""
import torch
n = 1 << 32  # 4 GiB
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(f, reps=10):
    for _ in range(3): f()
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best / 1e3
av = a.view(torch.int64)
bv = b.view(torch.int64)
print("fill  (write only): %.0f GB/s" % (n / t(lambda: av.fill_(7)) / 1e9))
print("sum   (read only) : %.0f GB/s" % (n / t(lambda: av.sum()) / 1e9))
print("copy  (read+write): %.0f GB/s" % (2 * n / t(lambda: bv.copy_(av)) / 1e9))
