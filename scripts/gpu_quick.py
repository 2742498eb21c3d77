"""Scratch GPU check: parity of the CUDA path vs the oracle on small batches + a first timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from bio_b200 import _cabi as cabi, synth

ctx = cabi.Context(0)

def check(name, mode, omode, bases, off, hint=0, **kw):
    p = cabi.make_params(mode, max_read_len=hint, **kw)
    try:
        res = ctx.run(p, bases, off)
    except Exception as e:
        print(f"{name}: EXC {e}", flush=True)
        return False
    okw = dict(kw); okw.pop('want_pos', None)
    ref = oracle.run_batch(bases, off, omode, threads=8, **okw)
    ok = (np.array_equal(res['off'], ref['off']) and np.array_equal(res['val'], ref['val'])
          and np.array_equal(res['status'], ref['status']))
    if ref['pos'] is not None and res['pos'] is not None:
        ok = ok and np.array_equal(res['pos'], ref['pos'])
    print(f"{name}: {'OK' if ok else 'MISMATCH'} n_out={res['total']} ref={len(ref['val'])} ties={ref['ties']}", flush=True)
    if not ok:
        n = min(len(res['val']), len(ref['val']))
        bad = np.nonzero(res['val'][:n] != ref['val'][:n])[0]
        print("  first val mismatch", bad[:5], "off eq", np.array_equal(res['off'], ref['off']),
              "status eq", np.array_equal(res['status'], ref['status']))
        if res['pos'] is not None and ref['pos'] is not None:
            badp = np.nonzero(res['pos'][:n] != ref['pos'][:n])[0]
            print("  first pos mismatch", badp[:5])
    return ok

allok = True
b, o = synth.uniform_reads(20000, 150, 42)
allok &= check("nthash 150bp hint", cabi.MODE_NTHASH, oracle.MODE_NTHASH, b, o, hint=150, k=21)
allok &= check("nthash 150bp nohint", cabi.MODE_NTHASH, oracle.MODE_NTHASH, b, o, k=21)
allok &= check("nthash fwd", cabi.MODE_NTHASH, oracle.MODE_NTHASH, b, o, hint=150, k=21, canonical=False)
allok &= check("minimizer 150bp hint", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b, o, hint=150, k=21, w=11)
allok &= check("minimizer 150bp nohint", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b, o, k=21, w=11)
allok &= check("minimizer w=1", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b, o, hint=150, k=21, w=1)
allok &= check("minimizer k5w3", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b, o, hint=150, k=5, w=3)
allok &= check("syncmer 150bp", cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, b, o, hint=150, k=21, s=11)
allok &= check("syncmer s=k", cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, b, o, hint=150, k=21, s=21)
allok &= check("minimizer circular", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b, o, hint=150, k=21, w=11, circular=True)
L = synth.ont_like_lengths(300, 44)
b2, o2 = synth.ragged_reads(L, 44)
allok &= check("syncmer ONT", cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, b2, o2, k=21, s=11)
allok &= check("minimizer ONT", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b2, o2, k=21, w=11)
allok &= check("nthash ONT", cabi.MODE_NTHASH, oracle.MODE_NTHASH, b2, o2, k=21)
allok &= check("minimizer ONT hint", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b2, o2, hint=int(L.max()), k=21, w=11)
lens = np.array([0, 5, 30, 31, 32, 150, 0, 0, 400, 20, 31, 1000, 3, 151] * 50)
b3, o3 = synth.ragged_reads(lens, 7, alphabet=b"ACGTNacgtRY")
allok &= check("minimizer ragged", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b3, o3, k=21, w=11)
allok &= check("syncmer ragged", cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, b3, o3, k=21, s=11)
b4, o4 = synth.ragged_reads([150] * 500, 9, alphabet=b"A")
allok &= check("minimizer polyA", cabi.MODE_MINIMIZER, oracle.MODE_MINIMIZER, b4, o4, hint=150, k=21, w=11)
b5, o5 = synth.ragged_reads([5000] * 20, 9, alphabet=b"AC")
allok &= check("syncmer lowcomplex", cabi.MODE_SYNCMER, oracle.MODE_SYNCMER, b5, o5, k=21, s=11)
for canon in (True, False):
    allok &= check(f"kmer 150bp canon={canon}", cabi.MODE_KMER, oracle.MODE_KMER, b, o, hint=150, k=21, canonical=canon)
    allok &= check(f"kmer ONT canon={canon}", cabi.MODE_KMER, oracle.MODE_KMER, b2, o2, k=31, canonical=canon)
    allok &= check(f"kmer ragged canon={canon}", cabi.MODE_KMER, oracle.MODE_KMER, b3, o3, k=5, canonical=canon)
    allok &= check(f"kmer circular canon={canon}", cabi.MODE_KMER, oracle.MODE_KMER, b, o, hint=150, k=21, canonical=canon, circular=True)
bb = b3.copy(); bb[::97] = ord('-')
for canon in (True, False):
    allok &= check(f"kmer illegal canon={canon}", cabi.MODE_KMER, oracle.MODE_KMER, bb, o3, k=7, canonical=canon)
for fr in (1, 2, 3, -1, -2, -3):
    allok &= check(f"protein 150bp frame {fr}", cabi.MODE_PROTEIN, oracle.MODE_PROTEIN, b, o, hint=150, k=11, frame=fr)
    allok &= check(f"protein ONT frame {fr}", cabi.MODE_PROTEIN, oracle.MODE_PROTEIN, b2, o2, k=11, frame=fr)
    allok &= check(f"protein ragged frame {fr}", cabi.MODE_PROTEIN, oracle.MODE_PROTEIN, b3, o3, k=5, frame=fr)
allok &= check("protein table 11 k=40", cabi.MODE_PROTEIN, oracle.MODE_PROTEIN, b2, o2, k=40, frame=-2, codon_table=11)
print("ALL OK" if allok else "SOME FAILED", flush=True)

dev = torch.device("cuda:0")
for n_reads in (10_000_000,):
    bases, off = synth.device_uniform_reads(n_reads, 150, 43, dev)
    nb = n_reads * 150
    for name, mode, kw in (("minimizer", cabi.MODE_MINIMIZER, dict(k=21, w=11)),
                           ("nthash", cabi.MODE_NTHASH, dict(k=21)),
                           ("syncmer", cabi.MODE_SYNCMER, dict(k=21, s=11)),
                           ("kmer", cabi.MODE_KMER, dict(k=21)),
                           ("protein", cabi.MODE_PROTEIN, dict(k=11, frame=1))):
        p = cabi.make_params(mode, max_read_len=150, **kw)
        cap = int(cabi.lib().b200sk_output_bound(p, nb, n_reads, 0))
        val = torch.empty(cap, dtype=torch.int64, device=dev)
        pos = torch.empty(cap, dtype=torch.int32, device=dev)
        ooff = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
        st = torch.empty(n_reads, dtype=torch.int32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        rc, total = ctx.run_device(p, bases, off, nb, val, pos, ooff, st)
        print(name, "rc", rc, "total", total, "per read", total / n_reads, flush=True)
        for _ in range(3):
            ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for _ in range(K):
            ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, st, flags)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{name}: {ms:.3f} ms/step  {nb / ms / 1e6:.1f} Gbases/s  flags={flags.item()}", flush=True)
        del val, pos, ooff, st
