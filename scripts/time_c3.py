"""Time C3-geometry launches (dev helper): python scripts/time_c3.py [reads] [w] ; prints one JSON line"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bio_b200 import _cabi as cabi, synth
dev = torch.device("cuda:0")
ctx = cabi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 11
RL = int(os.environ.get("READ_LEN", 150))
bases, off = synth.device_uniform_reads(n, RL, 43, dev)
nb = n * RL
p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=w, max_read_len=RL)
cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 0))
val = torch.empty(cap, dtype=torch.int64, device=dev)
pos = torch.empty(cap, dtype=torch.int32, device=dev)
ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
st = torch.empty(n, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
for _ in range(3):
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
chk = int(val[:total].sum().item()) ^ int(pos[:total].sum().item())
print(json.dumps({"walker": os.environ.get("B200SK_WALKER", "exact"), "spin_ns": os.environ.get("B200SK_SPIN_NS", "0"), "w": w, "read_len": RL,
                  "ms": round(ms, 4), "Gbases_per_s": round(nb / ms / 1e6, 1), "elements": total, "checksum": chk}), flush=True)
