"""ONT-like long reads (C4 geometry) through the device entry point: timing for one mode (dev helper / ncu target)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bio_b200 import _cabi as cabi, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "syncmer"
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
ctx = cabi.Context(0)
L = synth.ont_like_lengths(nr, 44)
o = np.zeros(nr + 1, dtype=np.int64); np.cumsum(L.astype(np.int64), out=o[1:])
nb = int(o[-1])
g = torch.Generator(device=dev); g.manual_seed(44)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
bases = torch.zeros(nb + 64, dtype=torch.uint8, device=dev)
bases[:nb] = lut[torch.randint(0, 4, (nb,), generator=g, device=dev)]
off = torch.from_numpy(o).to(dev)
p = (cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=int(L.max())) if mode == "syncmer"
     else cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=int(L.max())))
cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, nr, 0))
val = torch.empty(cap, dtype=torch.int64, device=dev); pos = torch.empty(cap, dtype=torch.int32, device=dev)
ooff = torch.empty(nr + 1, dtype=torch.int64, device=dev); st = torch.empty(nr, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
for _ in range(3):
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"mode": mode, "reads": nr, "bases": nb, "ms": ms, "bases_per_s": nb / ms * 1e3, "elements": total}))
