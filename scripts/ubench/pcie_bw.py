"""PCIe copy bandwidth with pinned host memory: D2H alone, H2D alone, both at once (dev helper)."""
import torch, time
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
def d2h():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
def both():
    d2h(); h2d()
print("D2H  %.1f GB/s" % (n / run(d2h) / 1e9))
print("H2D  %.1f GB/s" % (n / run(h2d) / 1e9))
t = run(both)
print("both %.1f GB/s each (%.1f total)" % (n / t / 1e9, 2 * n / t / 1e9))
