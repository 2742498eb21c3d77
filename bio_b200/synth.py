"""Seeded synthetic inputs shared by tests and bench: concatenated ASCII bases + offsets."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def uniform_reads(n_reads, read_len, seed):
    """n_reads x read_len i.i.d. uniform ACGT (C2/C3/C5 style).  Returns (bases u8, off u64)."""
    rng = np.random.default_rng(seed)
    bases = _ACGT[rng.integers(0, 4, size=n_reads * read_len, dtype=np.uint8)]
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return bases, off


def ragged_reads(lengths, seed, alphabet=b"ACGT"):
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.uint64)
    off = np.zeros(len(lengths) + 1, dtype=np.uint64)
    np.cumsum(lengths, out=off[1:])
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    bases = alpha[rng.integers(0, len(alpha), size=int(off[-1]))]
    return bases, off


def ont_like_lengths(n_reads, seed, mean=10000.0, sigma=0.6, lo=200, hi=200000):
    """Lognormal read lengths with the given mean (C4)."""
    rng = np.random.default_rng(seed)
    mu = np.log(mean) - 0.5 * sigma * sigma
    L = rng.lognormal(mu, sigma, size=n_reads)
    return np.clip(L, lo, hi).astype(np.uint64)


def device_uniform_reads(n_reads, read_len, seed, device):
    """Same distribution generated directly in HBM (torch is plumbing here); padded by 64 bytes so
    the 16-byte TMA units of the last tile stay inside the allocation."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = n_reads * read_len
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    buf = torch.empty(n + 64, dtype=torch.uint8, device=device)
    step = 1 << 27
    for s in range(0, n, step):
        e = min(n, s + step)
        idx = torch.randint(0, 4, (e - s,), generator=g, device=device, dtype=torch.int64)
        buf[s:e] = lut[idx]
    buf[n:] = 0
    off = torch.arange(n_reads + 1, dtype=torch.int64, device=device) * read_len
    return buf, off
