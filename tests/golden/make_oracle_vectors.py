"""Regression fixtures produced by the CPU oracle on fixed seeded inputs (tests/golden/oracle_vectors.npz).
The GPU parity tests compare the CUDA path against these committed arrays as well as against a live
oracle run, so a silent change of either side is caught.  Regenerate with: python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from bio_b200 import synth  # noqa: E402

CASES = {
    # name: (oracle mode, kwargs)
    "nthash_k21": (oracle.MODE_NTHASH, dict(k=21)),
    "nthash_k21_fwd": (oracle.MODE_NTHASH, dict(k=21, canonical=False)),
    "minimizer_k21_w11": (oracle.MODE_MINIMIZER, dict(k=21, w=11)),
    "minimizer_k5_w3": (oracle.MODE_MINIMIZER, dict(k=5, w=3)),
    "syncmer_k21_s11": (oracle.MODE_SYNCMER, dict(k=21, s=11)),
    "kmer_k21": (oracle.MODE_KMER, dict(k=21)),
    "kmer_k5_both": (oracle.MODE_KMER, dict(k=5, canonical=False)),
    "protein_k11_f1": (oracle.MODE_PROTEIN, dict(k=11, frame=1)),
    "protein_k11_fm2": (oracle.MODE_PROTEIN, dict(k=11, frame=-2)),
}


def inputs():
    b1, o1 = synth.uniform_reads(48, 150, 1234)
    lens = [0, 5, 30, 31, 32, 150, 0, 400, 20, 1000, 3, 151, 2500, 77]
    b2, o2 = synth.ragged_reads(lens, 99, alphabet=b"ACGTACGTACGTNacgtRYK")
    bases = np.concatenate([b1, b2])
    off = np.concatenate([o1, o2[1:] + o1[-1]])
    return bases, off


if __name__ == "__main__":
    bases, off = inputs()
    out = dict(bases=bases, off=off)
    for name, (mode, kw) in CASES.items():
        r = oracle.run_batch(bases, off, mode, threads=1, **kw)
        out[name + "/val"] = r["val"]
        out[name + "/pos"] = r["pos"]
        out[name + "/off"] = r["off"]
        out[name + "/status"] = r["status"]
        print(name, len(r["val"]), "ties", r["ties"])
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
