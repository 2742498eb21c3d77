"""A/B of the minimizer walkers (dev helper).  Run twice, B200SK_WALKER=keyed and unset (= the exact 64-bit window); each run checks parity against
the oracle on small batches (uniform, ragged, low-complexity, long reads) and times C3-geometry batches.
    python scripts/ab_walker.py [reads]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from bio_b200 import _cabi as cabi, synth

dev = torch.device("cuda:0")
ctx = cabi.Context(0)
tag = os.environ.get("B200SK_WALKER", "exact")
ok_all = True


def check(name, bases, off, hint=0, **kw):
    global ok_all
    p = cabi.make_params(cabi.MODE_MINIMIZER, max_read_len=hint, **kw)
    res = ctx.run(p, bases, off)
    ref = oracle.run_batch(bases, off, oracle.MODE_MINIMIZER, threads=8, **kw)
    ok = (np.array_equal(res["off"], ref["off"]) and np.array_equal(res["val"], ref["val"])
          and np.array_equal(res["status"], ref["status"]) and np.array_equal(res["pos"], ref["pos"]))
    ok_all &= ok
    print(f"[{tag}] {name}: {'OK' if ok else 'MISMATCH'} n_out={res['total']} ref={len(ref['val'])}", flush=True)
    if not ok:
        cg, cr = np.diff(res["off"].astype(np.int64)), np.diff(ref["off"].astype(np.int64))
        badr = np.nonzero(cg != cr)[0]
        print("   reads with a different count:", len(badr), badr[:8], "gpu-ref", (cg - cr)[badr[:8]])
        for r in badr[:2]:
            g = res["pos"][res["off"][r]:res["off"][r + 1]].tolist()
            f = ref["pos"][ref["off"][r]:ref["off"][r + 1]].tolist()
            d = next((i for i in range(min(len(g), len(f))) if g[i] != f[i]), min(len(g), len(f)))
            print("   read", r, "lane", r % 32, "first diff at entry", d, "gpu", g[max(0, d - 3):d + 4], "ref", f[max(0, d - 3):d + 4])
        n = min(len(res["val"]), len(ref["val"]))
        print("   first val mismatch", np.nonzero(res["val"][:n] != ref["val"][:n])[0][:5],
              "first pos mismatch", np.nonzero(res["pos"][:n] != ref["pos"][:n])[0][:5],
              "off eq", np.array_equal(res["off"], ref["off"]))


b, o = synth.uniform_reads(50000, 150, 42)
for w in (3, 5, 11, 15):
    check(f"uniform 150bp k21 w{w}", b, o, hint=150, k=21, w=w)
check("uniform 150bp k5 w3 (tie-heavy)", b, o, hint=150, k=5, w=3)
check("uniform 150bp k7 w11 (tie-heavy)", b, o, hint=150, k=7, w=11)
check("circular", b, o, hint=150, k=21, w=11, circular=True)
lens = np.array([0, 5, 30, 31, 32, 150, 0, 0, 400, 20, 31, 1000, 3, 151, 41, 42, 43] * 50)
b3, o3 = synth.ragged_reads(lens, 7, alphabet=b"ACGTNacgtRY")
check("ragged IUPAC", b3, o3, k=21, w=11)
check("ragged IUPAC w5", b3, o3, k=21, w=5)
b4, o4 = synth.ragged_reads([150] * 500, 9, alphabet=b"A")
check("polyA", b4, o4, hint=150, k=21, w=11)
b5, o5 = synth.ragged_reads([150] * 5000, 9, alphabet=b"AC")
check("AC low complexity", b5, o5, hint=150, k=21, w=11)
L = synth.ont_like_lengths(300, 44)
b2, o2 = synth.ragged_reads(L, 44)
check("ONT", b2, o2, k=21, w=11)
check("ONT hint", b2, o2, hint=int(L.max()), k=21, w=11)
b6, o6 = synth.ragged_reads([5000] * 20, 9, alphabet=b"AC")
check("ONT low complexity", b6, o6, k=21, w=11)
print(f"[{tag}] PARITY", "ALL OK" if ok_all else "FAILED", flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
bases, off = synth.device_uniform_reads(n, 150, 43, dev)
nb = n * 150
for w in (11, 5, 15):
    p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=w, max_read_len=150)
    cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 0))
    val = torch.empty(cap, dtype=torch.int64, device=dev)
    pos = torch.empty(cap, dtype=torch.int32, device=dev)
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
    for _ in range(5):
        ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    chk = int(val[:total].sum().item()) ^ int(pos[:total].sum().item())
    print(json.dumps({"walker": tag, "w": w, "reads": n, "ms": ms, "Gbases_per_s": nb / ms / 1e6, "elements": total,
                      "checksum": chk, "flags": int(flags.item())}), flush=True)
    del val, pos, ooff, st
