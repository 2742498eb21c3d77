"""Secondary configs (C2 ntHash, C4 syncmer on ONT-like reads, C5 protein 6 frames, k-mer codes): device-resident
throughput + bytes moved, printed as JSON lines (for DESIGN.md / profiles; the driver contract lives in bench.py)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bio_b200 import _cabi as cabi, synth

dev = torch.device("cuda:0")
ctx = cabi.Context(0)
PEAK = 6448.1


def run(name, p, bases, off, nb, n, steps=5, frames=None):
    cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 1 if p.mode in (0, 1, 4, 6) else 0))
    val = torch.empty(cap, dtype=torch.int64, device=dev)
    pos = torch.empty(cap, dtype=torch.int32, device=dev) if p.want_pos else None
    ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
    assert rc == 0, rc
    for _ in range(3):
        ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg = nb + 8 * n + total * (12 if p.want_pos else 8) + 8 * n
    print(json.dumps({"config": name, "ms": ms, "bases_per_s": nb / ms * 1e3, "elements": total,
                      "algorithmic_GBps": alg / ms / 1e6, "hbm_frac_of_measured": alg / ms / 1e6 / PEAK,
                      "bytes_per_base": alg / nb}), flush=True)
    del val, pos, ooff, st


n = int(os.environ.get("READS", 10_000_000))
bases, off = synth.device_uniform_reads(n, 150, 42, dev)
nb = n * 150
run("C2 ntHash k=21 canonical, 10M x 150bp (values only)", cabi.make_params(cabi.MODE_NTHASH, 21, max_read_len=150, want_pos=False), bases, off, nb, n)
run("C3-geometry minimizer k=21 w=11, 10M x 150bp", cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150), bases, off, nb, n)
run("syncmer k=21 s=11, 10M x 150bp", cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=150), bases, off, nb, n)
run("C1-geometry k-mer codes k=21 canonical, 10M x 150bp (values only)", cabi.make_params(cabi.MODE_KMER, 21, max_read_len=150, want_pos=False), bases, off, nb, n)
run("SimHash k=31 m=5 scale=5, 10M x 150bp (values only)", cabi.make_params(cabi.MODE_SIMHASH, 31, m=5, scale=5, max_read_len=150, want_pos=False), bases, off, nb, n)
run("ProteinMinimizer k=10 w=5 frame 1, 10M x 150bp", cabi.make_params(cabi.MODE_PROTEIN_MINIMIZER, 10, w=5, frame=1, max_read_len=150), bases, off, nb, n)
for fr in (1, 2, 3, -1, -2, -3):
    run(f"C5 protein k=11 frame {fr}, 10M x 150bp (values only)", cabi.make_params(cabi.MODE_PROTEIN, 11, frame=fr, max_read_len=150, want_pos=False), bases, off, nb, n)
del bases, off
# C4: ONT-like long reads
nr = int(os.environ.get("ONT_READS", 200_000))
L = synth.ont_like_lengths(nr, 44)
o = np.zeros(nr + 1, dtype=np.int64); np.cumsum(L.astype(np.int64), out=o[1:])
nb = int(o[-1])
g = torch.Generator(device=dev); g.manual_seed(44)
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
bases = torch.zeros(nb + 64, dtype=torch.uint8, device=dev)
bases[:nb] = lut[torch.randint(0, 4, (nb,), generator=g, device=dev)]
off = torch.from_numpy(o).to(dev)
run(f"C4 syncmer k=21 s=11, {nr} ONT-like reads mean 10kb", cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=int(L.max())), bases, off, nb, nr)
run(f"minimizer k=21 w=11, {nr} ONT-like reads mean 10kb", cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=int(L.max())), bases, off, nb, nr)
