"""PCIe probe (dev helper): pinned H2D / D2H bandwidth alone and concurrently, by chunk size -- the ceiling of the e2e path."""
import json, sys, torch
dev = torch.device("cuda:0")
out = {}
for mb in (64, 384, 1024):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = max(4, 8192 // mb)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    run(True, True)
    out["%d MiB" % mb] = {"h2d_alone_GBps": round(run(True, False), 1), "d2h_alone_GBps": round(run(False, True), 1),
                          "both_each_GBps": round(run(True, True), 1)}
print(json.dumps(out))
