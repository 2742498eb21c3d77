"""Record feeder measurement: FASTQ text of n x 150 bp reads in HBM -> records -> packed
bases + offsets (b200sk_fastx_parse_device), and text in pinned host memory -> minimizers (b200sk_run_fastx),
next to the oracle restatement of fastx.Reader.Read on one host core.  One JSON line per measurement."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from bio_b200 import _cabi as cabi


def fastq_text(n, L, seed):
    """n records '@r%09d\\n<L bases>\\n+\\n<L quals>\\n' as one uint8 array (vectorised)."""
    rng = np.random.default_rng(seed)
    rec = 1 + 10 + 1 + L + 1 + 2 + L + 1
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0] = ord("@")
    a[:, 1] = ord("r")
    idx = np.arange(n, dtype=np.int64)
    for d in range(9):
        a[:, 2 + d] = ord("0") + (idx // 10 ** (8 - d)) % 10
    a[:, 11] = 10
    a[:, 12:12 + L] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    a[:, 12 + L] = 10
    a[:, 13 + L] = ord("+")
    a[:, 14 + L] = 10
    a[:, 15 + L:15 + 2 * L] = rng.integers(33, 74, size=(n, L), dtype=np.uint8)
    a[:, 15 + 2 * L] = 10
    return a.reshape(-1)


def fasta_text(n_rec, rec_len, width, seed):
    """n_rec records '>s%09d\n' + rec_len bases in lines of `width` (vectorised; rec_len a multiple of width)."""
    rng = np.random.default_rng(seed)
    rows = rec_len // width
    rec = 12 + rows * (width + 1)
    a = np.empty((n_rec, rec), dtype=np.uint8)
    a[:, 0] = ord(">")
    a[:, 1] = ord("s")
    idx = np.arange(n_rec, dtype=np.int64)
    for d in range(9):
        a[:, 2 + d] = ord("0") + (idx // 10 ** (8 - d)) % 10
    a[:, 11] = 10
    body = a[:, 12:].reshape(n_rec, rows, width + 1)
    body[:, :, :width] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n_rec, rows, width), dtype=np.uint8)]
    body[:, :, width] = 10
    return a.reshape(-1)


def bench_fasta(ctx, peaks):
    n_rec, rec_len, width = 200_000, 9_960, 60  # ~2 GB of 60-column FASTA, 10 kb records
    text = fasta_text(n_rec, rec_len, width, 47)
    nb = text.size
    pad = np.zeros((nb + 15) // 16 * 16 + 16, dtype=np.uint8)
    pad[:nb] = text
    d_text = torch.from_numpy(pad).cuda()
    for _ in range(2):
        info = ctx.fastx_parse_device(d_text, nb, 0, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(5):
        e0.record()
        info = ctx.fastx_parse_device(d_text, nb, 0, True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    assert info.n_records == n_rec and info.n_bases == n_rec * rec_len
    t = float(np.median(ms)) / 1e3
    nlines = n_rec * (1 + rec_len // width)
    alg = nb + n_rec * rec_len + 2 * 8 * n_rec + 8 * nlines + 8 * nlines  # text in; bases, record tables, line table, out_pos out
    print(json.dumps({"config": f"feeder: FASTA {n_rec} x {rec_len} bp in {width}-column lines ({nb / 1e9:.2f} GB text) -> records in HBM",
                      "ms": t * 1e3, "text_GBps": nb / t / 1e9, "bases_per_s": n_rec * rec_len / t,
                      "algorithmic_GBps": alg / t / 1e9, "hbm_frac_of_measured": alg / t / 1e9 / peaks["hbm_gbs"]}), flush=True)
    del d_text


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    L = 150
    reps = 5
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    text = fastq_text(n, L, 46)
    nb = text.size
    ctx = cabi.Context(0)
    pad = np.zeros((nb + 15) // 16 * 16 + 16, dtype=np.uint8)
    pad[:nb] = text
    d_text = torch.from_numpy(pad).cuda()
    torch.cuda.synchronize()
    # ---- device-resident parse
    for _ in range(2):
        info = ctx.fastx_parse_device(d_text, nb, 0, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        e0.record()
        info = ctx.fastx_parse_device(d_text, nb, 0, True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    assert info.n_records == n and info.n_bases == n * L
    t = float(np.median(ms)) / 1e3
    # compulsory traffic: the text once in; bases + read_off + rec_off + qual_off + line table out
    alg = nb + n * L + 3 * 8 * n + 8 * 4 * n
    print(json.dumps({"config": f"feeder: FASTQ {n} x {L} bp ({nb / 1e9:.2f} GB text) -> records in HBM",
                      "ms": t * 1e3, "text_GBps": nb / t / 1e9, "bases_per_s": n * L / t, "records_per_s": n / t,
                      "algorithmic_GBps": alg / t / 1e9, "hbm_frac_of_measured": alg / t / 1e9 / peaks["hbm_gbs"],
                      "bytes_per_base": alg / (n * L)}), flush=True)
    bench_fasta(ctx, peaks)
    # ---- text in pinned host memory -> minimizers on the host (one C-ABI call)
    L_ = cabi.lib()
    hp = L_.b200sk_alloc_pinned(nb)
    import ctypes as C
    C.memmove(hp, text.ctypes.data, nb)
    htext = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb,))
    p = cabi.make_params(cabi.MODE_MINIMIZER, k=21, w=11, max_read_len=L, pos_width=1)
    res = ctx.run_fastx(p, htext, copy=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = ctx.run_fastx(p, htext, copy=False)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    print(json.dumps({"config": f"e2e: FASTQ text (pinned host) -> minimizers k=21 w=11 (host), {n} x {L} bp",
                      "ms": t * 1e3, "bases_per_s": n * L / t, "text_GBps": nb / t / 1e9,
                      "h2d_bytes": nb, "d2h_bytes": int(res["total"]) * 9 + 12 * n, "elements": int(res["total"])}),
          flush=True)
    total_one_shot = int(res["total"])
    chk_one_shot = int(res["val"][:1000].sum())
    # ---- the same through the pipelined reader (b200sk_fxstream: two slots, chunks of chunk_mb of text)
    for chunk_mb in (64, 256):
        stream = cabi.FastxStream(p, htext, chunk_bytes=chunk_mb << 20, copy=False)
        tot = sum(c["total"] for c in stream)  # warm-up: allocates both slots' buffers
        assert tot == total_one_shot
        ts = []
        for _ in range(3):
            stream.rewind()
            t0 = time.perf_counter()
            tot, nrec, nch, first = 0, 0, 0, None
            for c in stream:
                if first is None:
                    first = int(c["val"][:1000].sum())
                tot += c["total"]; nrec += int(c["info"].n_records); nch += 1
            ts.append(time.perf_counter() - t0)
        assert tot == total_one_shot and nrec == n and first == chk_one_shot
        t = float(np.median(ts))
        print(json.dumps({"config": f"e2e pipelined: FASTQ text (pinned host) -> minimizers k=21 w=11 (host), {n} x {L} bp, "
                                    f"b200sk_fxstream, {nch} chunks of {chunk_mb} MiB",
                          "ms": t * 1e3, "bases_per_s": n * L / t, "text_GBps": nb / t / 1e9,
                          "h2d_bytes": nb, "d2h_bytes": tot * 9 + 12 * n, "elements": tot}), flush=True)
        stream.close()
    # ---- CPU: oracle restatement of Reader.Read on one core, bounded sample
    m = min(n, 1_000_000)
    sample = text[:m * (nb // n)].tobytes()
    t0 = time.perf_counter()
    o = oracle.fastx_parse(sample)
    t = time.perf_counter() - t0
    assert o["n_records"] == m
    print(json.dumps({"config": f"CPU oracle (Reader.Read restatement, 1 core), {m} x {L} bp",
                      "ms": t * 1e3, "text_GBps": len(sample) / t / 1e9, "bases_per_s": m * L / t}), flush=True)
    L_.b200sk_free_pinned(hp)
    ctx.close()


if __name__ == "__main__":
    main()
