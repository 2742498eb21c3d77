"""One mode on 10M x 150bp device-resident reads (dev helper / ncu target): python scripts/run_mode.py nthash|kmer|protein|simhash [steps]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bio_b200 import _cabi as cabi, synth
mode = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = int(os.environ.get("READS", 10_000_000))
dev = torch.device("cuda:0"); ctx = cabi.Context(0)
bases, off = synth.device_uniform_reads(n, 150, 42, dev); nb = n * 150
p = {"nthash": cabi.make_params(cabi.MODE_NTHASH, 21, max_read_len=150, want_pos=False),
     "kmer": cabi.make_params(cabi.MODE_KMER, 21, max_read_len=150, want_pos=False),
     "protein": cabi.make_params(cabi.MODE_PROTEIN, 11, frame=1, max_read_len=150, want_pos=False),
     "simhash": cabi.make_params(cabi.MODE_SIMHASH, 31, m=5, scale=5, max_read_len=150, want_pos=False),
     "protmin": cabi.make_params(cabi.MODE_PROTEIN_MINIMIZER, 10, w=5, frame=1, max_read_len=150),
     "minimizer": cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150),
     "syncmer": cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=150)}[mode]
cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 1 if p.mode in (0, 1, 4, 6) else 0))
val = torch.empty(cap, dtype=torch.int64, device=dev)
pos = torch.empty(cap, dtype=torch.int32, device=dev) if p.want_pos else None
ooff = torch.empty(n + 1, dtype=torch.int64, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
for _ in range(3): ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
alg = nb + 16 * n + total * (12 if p.want_pos else 8)
print(json.dumps({"mode": mode, "ms": ms, "bases_per_s": nb / ms * 1e3, "elements": total, "alg_GBps": alg / ms / 1e6,
                  "hbm_frac": alg / ms / 1e6 / 6448.1}))
