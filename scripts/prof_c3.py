"""One C3-geometry launch for ncu (dev helper): python scripts/prof_c3.py [reads] [w]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bio_b200 import _cabi as cabi, synth
dev = torch.device("cuda:0")
ctx = cabi.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 11
bases, off = synth.device_uniform_reads(n, 150, 43, dev)
nb = n * 150
p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=w, max_read_len=150)
cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 0))
val = torch.empty(cap, dtype=torch.int64, device=dev)
pos = torch.empty(cap, dtype=torch.int32, device=dev)
ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
st = torch.empty(n, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
for _ in range(2):
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
torch.cuda.synchronize()
print("done", total)
