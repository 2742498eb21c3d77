#!/usr/bin/env python
"""bench.py -- bases/sec sketched (k=21, w=11 minimizer) on synthetic 150 bp reads, 1..8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the sketching hot path (sketches.NewMinimizerSketch + NextMinimizer/Index over every read,
reference: sketches/sketch.go:85,205) over ONE batch of 100 M synthetic reads (BASELINE.json config 3), through the
C ABI of libb200sketch.so.  With N GPUs the batch is SHARDED over the ranks ("scaling": "strong") and the gather of
the per-GPU arrays to rank 0 is INSIDE the step: the batch is dealt out in chunks of 32 768 reads (chunk c -> rank c % N), the
ranks' sketching kernels share ONE ordered output chain (b200sk_enqueue_device_sharded: the look-back over per-tile status
words runs across the GPUs) and every flush stores straight into rank 0's arrays over NVLink at its exact place.

JSON line (rank 0):
  value         whole-job bases/s, batch resident in HBM, gathered array complete on rank 0 when the clock stops
  e2e           same metric through the host entry point b200sk_run: pinned HOST buffers in, H2D and D2H inside
  roofline      algorithmic bytes of the sketching kernel / its CUDA-event duration vs the measured HBM copy peak
  parity        sampled oracle check of every config's device output, done BEFORE anything is timed (abort on mismatch)
  secondary     BASELINE.json configs 2, 4, 5 with their own roofline fractions
  cpu_baseline  the oracle (C restatement of the Go loops -- Go is absent here) on the host cores (N=1 only)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W, S, READ_LEN = 21, 11, 11, 150
SEED = 43
BLOCK = 1 << 19   # reads per generator block (one seed each)
CHUNK = int(os.environ.get("B200SK_BENCH_CHUNK", 1 << 15))   # reads per chunk of the sharded C3 batch: chunk c belongs to rank c % N (1024 tiles of 32 reads)
METRIC = "bases/sec sketched (k=21,w=11 minimizer)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads in the C3 batch (whole job)")
    ap.add_argument("--c2-reads", type=int, default=10_000_000)
    ap.add_argument("--c4-reads", type=int, default=1_000_000)
    ap.add_argument("--c5-reads", type=int, default=10_000_000)
    ap.add_argument("--parity-reads", type=int, default=100_000, help="reads per config in the oracle sample")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per GPU per e2e step (0 = the rank's shard, memory permitting)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=2_000_000, help="reads in the bounded CPU sample")
    ap.add_argument("--scale", type=int, default=100, help="FracMinHash scale of the reduced variants (f4)")
    ap.add_argument("--gather-pos", type=int, default=1, help="N > 1: also gather the uint8 positions on rank 0")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-reduce", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="e2e: do not bind the rank to its GPU's NUMA node")
    return ap.parse_args()


# ---------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the sketching kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def algorithmic_bytes(n_reads, n_bases, n_out, pos_bytes=4):
    # each base read once (1 B ASCII), 8 B read offset in; 8 B value (+ 4 B position) per
    # emitted element and 8 B output offset per read out.
    return n_bases + 8 * n_reads + (8 + pos_bytes) * n_out + 8 * n_reads


def shard_bounds(n, world):
    """equal-count contiguous shards, boundaries on generator blocks when n allows it"""
    nb = (n + BLOCK - 1) // BLOCK
    return [min(n, (nb * r // world) * BLOCK) if r < world else n for r in range(world + 1)]


def gen_uniform_block(torch, b, n_reads, read_len, seed, dev, lut, g):
    """reads [b * BLOCK, min((b + 1) * BLOCK, n_reads)) of the job-wide batch, drawn from the block's own seed"""
    g.manual_seed(seed * 1_000_003 + b)
    full = torch.randint(0, 4, (BLOCK * read_len,), generator=g, device=dev, dtype=torch.int64)
    cnt = (min(n_reads, (b + 1) * BLOCK) - b * BLOCK) * read_len
    return lut[full[:cnt]]


def gen_uniform_shard(torch, r0, r1, read_len, seed, dev, n_reads=None):
    """reads [r0, r1) of the job-wide batch (r0 on a block boundary): the union over the ranks is the same batch for
    every N.  Padded by 64 bytes (the 16-byte TMA units of the last tile)."""
    assert r0 % BLOCK == 0 or r0 == r1
    n = (r1 - r0) * read_len
    buf = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    buf[n:] = 0
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    pos = 0
    for b in range(r0 // BLOCK, (r1 + BLOCK - 1) // BLOCK):
        blk = gen_uniform_block(torch, b, r1, read_len, seed, dev, lut, g)
        buf[pos:pos + blk.numel()] = blk
        pos += blk.numel()
    off = torch.arange(r1 - r0 + 1, dtype=torch.int64, device=dev) * read_len
    return buf, off


def gen_uniform_chunks(torch, n_reads, rank, world, read_len, seed, dev):
    """this rank's part of the job-wide batch dealt out in chunks of CHUNK reads (chunk c -> rank c % world), back
    to back in chunk order.  world == 1: the whole batch, identical to gen_uniform_shard(0, n_reads)."""
    n_chunks = (n_reads + CHUNK - 1) // CHUNK
    mine = range(rank, n_chunks, world)
    n_local = sum(min(n_reads, (c + 1) * CHUNK) - c * CHUNK for c in mine)
    buf = torch.empty(n_local * read_len + 64, dtype=torch.uint8, device=dev)
    buf[n_local * read_len:] = 0
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    pos = 0
    per = BLOCK // CHUNK
    for b in range((n_reads + BLOCK - 1) // BLOCK):
        blk = gen_uniform_block(torch, b, n_reads, read_len, seed, dev, lut, g)
        for c in range(b * per, min(n_chunks, (b + 1) * per)):
            if c % world != rank:
                continue
            lo = (c * CHUNK - b * BLOCK) * read_len
            cnt = (min(n_reads, (c + 1) * CHUNK) - c * CHUNK) * read_len
            buf[pos:pos + cnt] = blk[lo:lo + cnt]
            pos += cnt
    off = torch.arange(n_local + 1, dtype=torch.int64, device=dev) * read_len
    return buf, off, n_local


def gather_ranges(torch, src, starts, lens, chunk=50_000_000):
    """concatenate src[starts[i] : starts[i] + lens[i]] on the device, in chunks of about `chunk` elements"""
    outs = []
    n = starts.numel()
    i = 0
    csum = torch.cumsum(lens, 0)
    while i < n:
        base = int(csum[i - 1].item()) if i else 0
        j = int(torch.searchsorted(csum, torch.tensor([base + chunk], device=csum.device)).item()) + 1
        j = min(max(j, i + 1), n)
        ls = lens[i:j]
        tot = int(ls.sum().item())
        if tot:
            ex = torch.cumsum(ls, 0) - ls
            idx = torch.repeat_interleave(starts[i:j] - ex, ls) + torch.arange(tot, device=src.device)
            outs.append(src[idx])
        i = j
    return torch.cat(outs) if outs else src[:0]


def parity_sample(torch, np, oracle, name, omode, okw, d_bases, d_off, val, pos, ooff, status, n_sample, seed, threads,
                  out_base=0):
    """Compare the device output of a random sample of reads with the oracle.  Returns a dict; raises on a mismatch."""
    n = d_off.numel() - 1
    g = torch.Generator(device=d_off.device)
    g.manual_seed(seed)
    ns = min(n_sample, n)
    idx = torch.unique(torch.randint(0, n, (ns,), generator=g, device=d_off.device))
    starts, lens = d_off[idx], d_off[idx + 1] - d_off[idx]
    hb = gather_ranges(torch, d_bases, starts, lens).cpu().numpy()
    ho = np.zeros(idx.numel() + 1, dtype=np.uint64)
    ho[1:] = torch.cumsum(lens, 0).cpu().numpy()
    ref = oracle.run_batch(hb, ho, omode, threads=threads, want_pos=pos is not None, **okw)
    ostart = ooff[idx] - out_base
    ocnt = ooff[idx + 1] - ooff[idx]
    gv = gather_ranges(torch, val, ostart, ocnt).cpu().numpy().view(np.uint64)
    mism = 0
    if not np.array_equal(ocnt.cpu().numpy().astype(np.uint64), np.diff(ref["off"]).astype(np.uint64)):
        mism += 1
    elif not np.array_equal(gv, ref["val"]):
        mism += int(np.count_nonzero(gv != ref["val"]))
    if pos is not None and mism == 0:
        gp = gather_ranges(torch, pos, ostart, ocnt).cpu().numpy().view(np.uint32)
        mism += int(np.count_nonzero(gp != ref["pos"]))
    if status is not None and not np.array_equal(status[idx].cpu().numpy(), ref["status"]):
        mism += 1
    return {"config": name, "reads_checked": int(idx.numel()), "elements_checked": int(len(ref["val"])),
            "mismatches": mism, "first_window_ties": int(ref["ties"])}


def timed(torch, fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


# ---------------------------------------------------------------- reference arm (CPU)
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  The reference is pure Go and neither a Go toolchain nor
    its four un-vendored arithmetic modules exist in this image, so this arm times the oracle: the line-by-line C
    restatement of sketches/sketch.go:205-309 (+ ntHash), one thread per host core over contiguous read shards --
    kind "port"."""
    if rank != 0:
        return
    import oracle
    from bio_b200 import synth
    cores = os.cpu_count() or 1
    n = args.cpu_reads
    bases, off = synth.uniform_reads(n, READ_LEN, SEED)
    nb = n * READ_LEN

    def step():
        oracle.run_batch(bases, off, oracle.MODE_MINIMIZER, k=K, w=W, threads=cores, want_output=False)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = nb * args.steps / dt
    sample = (f"{n} x {READ_LEN} bp uniform ACGT reads per step (seed {SEED}), same k/w; the full workload is one batch "
              f"of {args.reads} reads")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"C3 minimizer k={K} w={W}, {READ_LEN} bp reads, CPU sample", "reads_per_step": n},
        "cpu_baseline": {"value": v, "unit": "bases/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the Go algorithm (oracle/), not the Go binary: no Go toolchain in the "
                                 "image; published Go figure 18.3 Mbases/s/core (k=31,w=15, Ryzen 2700X).  The port "
                                 "allocates per read (3 malloc + one 150-byte copy) where Go pools its iterators; it "
                                 "still runs faster per core than the published Go figure"},
        "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import oracle  # the checker of the parity gate and the cpu_baseline leg; never on the measured path
    from bio_b200 import _cabi as cabi, synth

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ctx = cabi.Context(local_rank)
    L = cabi.lib()
    cores = max(1, (os.cpu_count() or 1) // world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_i64(x):
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    peak, peak_src = measured_peak()
    parity = []

    # ------------------------------------------------------------ C3: the headline batch, sharded over the ranks
    bases, off, n = gen_uniform_chunks(torch, args.reads, rank, world, READ_LEN, SEED, dev)
    nb = n * READ_LEN
    p = cabi.make_params(cabi.MODE_MINIMIZER, K, w=W, max_read_len=READ_LEN)
    cap = int(L.b200sk_output_bound(ctypes.byref(p), nb, n, 0))
    val = torch.empty(cap, dtype=torch.int64, device=dev)
    pos = torch.empty(cap, dtype=torch.int32, device=dev)
    ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    rc, n_out = ctx.run_device(p, bases, off, nb, val, pos, ooff, status)  # also sizes the scratch
    if rc != 0:
        raise RuntimeError("capacity estimate too small: need %d" % n_out)

    # parity gate: a random sample of this rank's reads against the oracle, before anything is timed
    pr = parity_sample(torch, np, oracle, "C3 minimizer k=21 w=11", oracle.MODE_MINIMIZER, dict(k=K, w=W), bases, off, val,
                       pos, ooff, status, max(1, args.parity_reads // world), 1000 + rank, cores)
    for key in ("reads_checked", "elements_checked", "mismatches", "first_window_ties"):
        pr[key] = allsum_i64(pr[key])
    parity.append(pr)
    if pr["mismatches"]:
        raise RuntimeError("PARITY FAILED on C3: %r" % pr)
    local_sum = int(val[:n_out].sum().item())  # wraps mod 2^64: a checksum of the value stream
    total_out = allsum_i64(n_out)
    checksum = allsum_i64(local_sum) & 0xFFFFFFFFFFFFFFFF

    # device-resident, local outputs: the per-GPU kernel number (what round 1 reported as `value`)
    ctx.timing_enable(True)
    res_ms = timed(torch, lambda: ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags),
                   args.steps, max(args.warmup, 3), barrier)
    kern_ms_sum, kern_n = ctx.timing_collect()
    ctx.timing_enable(False)
    kern_ms = kern_ms_sum / max(kern_n, 1)
    res_ms_max = allmax(res_ms)
    if int(flags.item()) != 0:
        raise RuntimeError("kernel flags %d" % int(flags.item()))

    # the step: shard -> sketch -> gathered uint64 array on rank 0
    gather = None
    sampler = ClockSampler(local_rank)
    if world == 1:
        if rank == 0:
            sampler.start()
        l0 = ctx.kernel_launches()
        ms_step = timed(torch, lambda: ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags),
                        args.steps, max(args.warmup, 3), barrier)
        launches = ctx.kernel_launches() - l0
        clocks = sampler.stop()
    else:
        # ONE output chain across the GPUs (b200sk_enqueue_device_sharded): the root owns the result arrays, every rank
        # owns a copy of the per-tile status words; all of them are CUDA-IPC buffers mapped into every process
        cap_total = int(L.b200sk_output_bound(ctypes.byref(p), args.reads * READ_LEN, args.reads, 0))
        n_tiles = (args.reads + 31) // 32
        mine = {"state": ctx.gather_create(n_tiles + 1)}
        if rank == 0:
            mine.update(val=ctx.gather_create(cap_total), pos=ctx.gather_create(cap_total // 8 + 8),
                        off=ctx.gather_create(args.reads + 1), status=ctx.gather_create(args.reads // 2 + 8))
        everyone = [None] * world
        dist.all_gather_object(everyone, {k: v[0] for k, v in mine.items()})
        opened = []

        def mapped(r, key):
            if r == rank:
                return mine[key][1]
            a = ctx.gather_open(everyone[r][key])
            opened.append(a)
            return a

        g_val, g_pos, g_off, g_status = (mapped(0, k) for k in ("val", "pos", "off", "status"))
        states = [mapped(r, "state") for r in range(world)]
        ps = cabi.make_params(cabi.MODE_MINIMIZER, K, w=W, max_read_len=READ_LEN, pos_width=1)
        ps_nopos = cabi.make_params(cabi.MODE_MINIMIZER, K, w=W, max_read_len=READ_LEN, want_pos=False)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        epoch = [0]

        def step(with_pos=bool(args.gather_pos)):
            epoch[0] += 1
            spec = cabi.ShardSpec()
            spec.rank, spec.n_ranks, spec.chunk_reads, spec.n_reads_global = rank, world, CHUNK, args.reads
            spec.epoch = epoch[0] % 16383 + 1
            for r in range(world):
                spec.state[r] = states[r]
            # every rank's flush stores straight into rank 0's arrays at its exact place (peer st.global over NVLink)
            ctx.enqueue_device_sharded(ps if with_pos else ps_nopos, spec, bases, off, nb, g_val, g_pos if with_pos else 0,
                                       g_off, g_status, cap_total, flags)
            dist.all_reduce(token)  # stream-ordered: no rank starts step e+1 before every rank has finished step e

        if rank == 0:
            sampler.start()
        l0 = ctx.kernel_launches()
        ms_step = allmax(timed(torch, step, args.steps, max(args.warmup, 3), barrier))
        launches = ctx.kernel_launches() - l0
        clocks = sampler.stop() if rank == 0 else None
        if int(flags.item()) != 0:
            raise RuntimeError("kernel flags %d" % int(flags.item()))
        gsum_ok = None
        if rank == 0:
            # the root must hold what one GPU would have produced: same count, same checksum, a monotone offset table
            goff = _as_tensor(torch, g_off, args.reads + 1, dev)
            tot = int(goff[args.reads].item())
            g = _as_tensor(torch, g_val, tot, dev)
            gsum_ok = bool(tot == total_out and (int(g.sum().item()) & 0xFFFFFFFFFFFFFFFF) == checksum
                           and bool((goff[1:] >= goff[:-1]).all().item()) and int(goff[0].item()) == 0)
            del g, goff
            if not gsum_ok:
                raise RuntimeError("PARITY FAILED: the gathered arrays differ from the per-rank results")
        barrier()
        # the same step gathering the uint64 arrays (+ offsets, statuses) only -- what BASELINE.json's north star names
        vo_ms = allmax(timed(torch, lambda: step(False), max(args.steps // 2, 3), 1, barrier))
        # diagnostic: the same chain with every rank storing into arrays of its OWN (same global indices, nothing but
        # the status words crosses NVLink) -- what the coupling of the ranks costs without the gather's traffic
        l_val = torch.empty(cap_total, dtype=torch.int64, device=dev)
        l_off = torch.empty(args.reads + 1, dtype=torch.int64, device=dev)
        l_st = torch.empty(args.reads, dtype=torch.int32, device=dev)

        def step_local():
            epoch[0] += 1
            spec = cabi.ShardSpec()
            spec.rank, spec.n_ranks, spec.chunk_reads, spec.n_reads_global = rank, world, CHUNK, args.reads
            spec.epoch = epoch[0] % 16383 + 1
            for r in range(world):
                spec.state[r] = states[r]
            ctx.enqueue_device_sharded(ps_nopos, spec, bases, off, nb, l_val.data_ptr(), 0, l_off.data_ptr(),
                                       l_st.data_ptr(), cap_total, flags)
            dist.all_reduce(token)

        chain_ms = allmax(timed(torch, step_local, max(args.steps // 2, 3), 1, barrier))
        del l_val, l_off, l_st
        nvbytes = (total_out - n_out) * (9 if args.gather_pos else 8) + (args.reads - n) * 12 if rank == 0 else 0
        nvbytes = allsum_i64(nvbytes)
        gather = {"how": "one output chain across the GPUs: the sketching kernels' look-back runs through every rank's copy "
                         "of the status words (posted 8-byte peer stores), each flush stores at its exact place in rank 0's "
                         "arrays (peer st.global over NVLink); no staging copy, no compaction, no counts exchanged; a "
                         "4-byte NCCL all-reduce separates the steps",
                  "gathered": "uint64 values" + (" + uint8 positions" if args.gather_pos else "") + " + per-read offsets and statuses, all on rank 0",
                  "bytes_into_rank0_over_nvlink": nvbytes, "ingress_GBps": nvbytes / (ms_step * 1e-3) / 1e9,
                  "kernels_local_ms": res_ms_max, "gathered_checksum_ok": gsum_ok, "chunk_reads": CHUNK,
                  "chain_only_ms": chain_ms,
                  "values_only": {"ms_per_step": vo_ms, "value": args.reads * READ_LEN / (vo_ms * 1e-3), "unit": "bases/s",
                                  "what": "the same step without the uint8 positions (uint64 arrays + offsets + statuses)"}}
        for a in opened:
            ctx.gather_close(a, False)
        barrier()
        for v in mine.values():
            ctx.gather_close(v[1], True)
    value = args.reads * READ_LEN / (ms_step * 1e-3)

    # f4 on the resident arrays: FracMinHash fraction + sort + unique per rank, the per-rank sketches gathered on
    # rank 0 (tiny now) and merged there.  Reported beside the full-stream step, not instead of it.
    reduced = None
    if not args.no_reduce:
        reduced = run_reduced_device(args, ctx, cabi, oracle, torch, np, dist, dev, rank, world, p, bases, off, nb, n, val,
                                     pos, ooff, status, flags, n_out, barrier, allmax, allsum_i64, cores)

    # roofline of the dominant (sketching) kernel, this rank's shard with local stores
    alg = algorithmic_bytes(n, nb, n_out)
    achieved = alg / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    if traffic and int(traffic.get("reads_per_launch", -1)) != n:
        traffic = None  # the committed capture is for another launch size
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "kernel": "k_sparse_warp<MINIMIZER,W=11>", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg,
                "bytes_per_base": alg / nb, "peak_source": peak_src,
                "kernel_share_of_step": kern_ms / ms_step,
                "binding_unit": "integer ALU pipe (profiles/, DESIGN.md 5.1), not HBM"}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")
    resident = {"value": args.reads * READ_LEN / (res_ms_max * 1e-3), "unit": "bases/s", "ms": res_ms_max,
                "what": "every rank sketches its shard into its own HBM, no gather (max over ranks)"}

    # end-to-end through the host entry point
    e2e = None
    if not args.no_e2e:
        del val, pos, ooff, status
        torch.cuda.empty_cache()
        e2e = run_e2e(args, ctx, cabi, bases, n, rank, world, dev, dist, torch, np, oracle, cores)
    del bases, off
    torch.cuda.empty_cache()

    secondary = []
    if not args.no_secondary:
        secondary = run_secondary(args, ctx, cabi, synth, oracle, rank, world, dev, dist, torch, np, peak, parity, cores,
                                  barrier, allmax, allsum_i64)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C3: NewMinimizerSketch k={K} w={W} over ONE batch of {args.reads} x {READ_LEN} bp "
                                   f"uniform-ACGT reads sharded over {world} GPU(s), uint64 arrays gathered on rank 0",
                       "reads": args.reads, "reads_per_gpu": n, "read_len": READ_LEN, "k": K, "w": W,
                       "minimizers_per_read": total_out / args.reads, "value_checksum": checksum,
                       "parallelism": f"records sharded over {world} GPU(s); gather inside the step",
                       "l2": "inputs (15 GB) and outputs (27 GB) per step far exceed the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "parity": {"all_ok": all(x["mismatches"] == 0 for x in parity), "configs": parity,
                       "note": "first_window_ties = sampled reads whose first window holds equal hashes (the only place "
                               "the unpinned sorts.Quicksort tie order could matter); C5/ProteinIterator: WYHASH_UNPINNED"},
            "per_gpu_resident": resident, "secondary": secondary, "reduced": reduced,
        }
        if gather:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reduced_device(args, ctx, cabi, oracle, torch, np, dist, dev, rank, world, p, bases, off, nb, n, val, pos, ooff,
                       status, flags, n_out, barrier, allmax, allsum_i64, cores):
    """sketch -> keep h <= MaxUint64/scale -> sort -> unique on every rank; per-rank sketches to rank 0; merge there."""
    scale = args.scale
    state = {"out": torch.empty(min(n_out + 1, n_out // max(scale, 1) * 2 * (W + 1) + (1 << 20)), dtype=torch.int64, device=dev)}
    res = {}
    gbuf = None
    if world > 1:
        capr = state["out"].numel()
        hbox = [None]
        if rank == 0:
            handle, gaddr = ctx.gather_create(capr * world)
            hbox = [handle]
        dist.broadcast_object_list(hbox, src=0)
        if rank != 0:
            gaddr = ctx.gather_open(hbox[0])
        gbuf = _as_tensor(torch, gaddr, capr * world, dev)
        merged = torch.empty(capr * world, dtype=torch.int64, device=dev) if rank == 0 else None
        counts_d = torch.zeros(world, dtype=torch.int64, device=dev)

    def step():
        ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags)
        rc, m = ctx.reduce_device(val, n_out, state["out"], scale=scale, unique=True)
        if rc != 0:  # the kept fraction of window minima is ~(w+1)/scale, not 1/scale: size from the count reported
            state["out"] = torch.empty(m + m // 8 + 1024, dtype=torch.int64, device=dev)
            ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags)
            rc, m = ctx.reduce_device(val, n_out, state["out"], scale=scale, unique=True)
            if rc != 0:
                raise RuntimeError("reduce: capacity")
        out = state["out"]
        state["m"] = m
        if world > 1:
            gbuf[rank * capr: rank * capr + m].copy_(out[:m])           # peer store of the reduced sketch (rank > 0)
            dist.all_gather_into_tensor(counts_d, torch.tensor([m], dtype=torch.int64, device=dev))
            if rank == 0:
                c = counts_d.cpu().numpy().astype(np.uint64)
                ctx.compact_segments(gaddr, np.arange(world, dtype=np.uint64) * np.uint64(capr), c,
                                     torch.cuda.current_stream(dev).cuda_stream)
                rc, mm = ctx.reduce_device(gbuf, int(c.sum()), merged, scale=1, unique=True)
                state["merged"] = mm

    ms = allmax(timed(torch, step, max(2, args.steps // 3), 1, barrier))
    m = state["m"]
    # parity of the reduction itself: the oracle's stream for the first reads of this rank, reduced with numpy
    ns = min(n, 50000)
    hb = bases[:ns * READ_LEN].cpu().numpy()
    ho = np.arange(ns + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    ref = oracle.run_batch(hb, ho, oracle.MODE_MINIMIZER, k=K, w=W, threads=cores)
    want = np.unique(ref["val"][ref["val"] <= np.uint64(((1 << 64) - 1) // max(scale, 1))])
    ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags)
    n_first = int(ooff[ns].item())
    small = torch.empty(len(want) + 1024, dtype=torch.int64, device=dev)
    rc, mm = ctx.reduce_device(val[:n_first].clone(), n_first, small, scale=scale, unique=True)
    ok = bool(rc == 0 and mm == len(want) and np.array_equal(small[:mm].cpu().numpy().view(np.uint64), want))
    if not ok:
        raise RuntimeError("PARITY FAILED on the reduced sketch (f4)")
    kept_sum, nv = allsum_i64(m), allsum_i64(0 if rank == 0 else m) * 8
    res = {"scale": scale, "what": "sketch + keep h <= MaxUint64/scale + radix sort + unique on every rank"
           + ("; reduced sketches stored into rank 0's buffer, merged (sort + unique) there" if world > 1 else ""),
           "value": args.reads * READ_LEN / (ms * 1e-3), "unit": "bases/s", "ms_per_step": ms,
           "elements_in": allsum_i64(n_out), "distinct_kept_per_rank_sum": kept_sum,
           "oracle_check": {"reads": ns, "ok": ok}}
    if world > 1:
        res["merged_distinct"] = state.get("merged") if rank == 0 else None
        res["bytes_over_nvlink"] = nv
        del gbuf
        ctx.gather_close(gaddr, rank == 0)
    return res


def _as_tensor(torch, addr, n, dev):
    """int64 view of n elements of raw device memory at addr (the gather buffer is library-owned)."""
    class _Mem:
        pass
    m = _Mem()
    m.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (addr, False), "version": 2}
    return torch.as_tensor(m, device=dev)


def run_e2e(args, ctx, cabi, d_bases, n_shard, rank, world, dev, dist, torch, np, oracle, cores):
    """Same metric through b200sk_run: inputs in pinned HOST memory, every step copies them to the device,
    sketches, and copies values, positions, offsets and statuses back to pinned host memory."""
    from bio_b200 import shard
    # the pinned buffers allocated below land on the NUMA node of this rank's GPU
    prev_affinity, binding = shard.bind_host_to_device(dev.index) if not args.no_bind else (None, "off")
    n = args.e2e_reads
    if n <= 0:
        try:
            avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
        except Exception:
            avail = 64 << 30
        per_read = READ_LEN + 8 + 23 * 12 * 1.3 + 12
        n = int(min(n_shard, 0.45 * avail / world / per_read))
    nb = n * READ_LEN
    L = cabi.lib()
    hb_ptr = L.b200sk_alloc_pinned(nb + 64)
    ho_ptr = L.b200sk_alloc_pinned((n + 1) * 8)
    if not hb_ptr or not ho_ptr:
        raise RuntimeError("pinned allocation failed")
    hb = np.ctypeslib.as_array(ctypes.cast(hb_ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nb,))
    ho = np.ctypeslib.as_array(ctypes.cast(ho_ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,))
    # fill the pinned batch buffer with the same synthetic reads (copied down from the device generator)
    torch.from_numpy(hb).copy_(d_bases[:nb])
    ho[:] = np.arange(n + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    torch.cuda.synchronize()
    # Index() values of a 150-bp read fit one byte: ask for uint8 positions (a quarter less D2H traffic)
    p = cabi.make_params(cabi.MODE_MINIMIZER, K, w=W, max_read_len=READ_LEN, pos_width=1)

    def step():
        return ctx.run(p, hb, ho, copy=False)

    res = step()  # warm-up: allocates the library's device + pinned output buffers
    res = step()
    n_out = res["total"]
    # the host path's output against the oracle on the first reads of the batch
    ns = min(n, 20000)
    ref = oracle.run_batch(hb[:ns * READ_LEN], ho[:ns + 1], oracle.MODE_MINIMIZER, k=K, w=W, threads=cores)
    m = int(ref["off"][-1])
    ok = bool(np.array_equal(res["val"][:m], ref["val"]) and np.array_equal(res["pos"][:m], ref["pos"].astype(np.uint8))
              and np.array_equal(res["off"][:ns + 1], ref["off"]))
    if not ok:
        raise RuntimeError("PARITY FAILED on the host path (b200sk_run)")
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        res = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    tot = torch.tensor([nb], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    out = {"value": int(tot.item()) * args.e2e_steps / dt, "unit": "bases/s",
           "h2d_bytes_per_step": nb + (n + 1) * 8,
           "d2h_bytes_per_step": n_out * 9 + (n + 1) * 8 + n * 4, "pos_width": 1,
           "reads_per_gpu_per_step": n, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
           "api": "b200sk_run (host pointers, pinned)", "oracle_check_first_reads": ns, "host_binding": binding}
    if not args.no_reduce:
        # the same batch through b200sk_run_reduced: only the FracMinHash sketch (sorted, distinct) comes back
        red = ctx.run_reduced(p, hb, ho, scale=args.scale, unique=True, copy=False)
        want = np.unique(ref["val"][ref["val"] <= np.uint64(((1 << 64) - 1) // max(args.scale, 1))])
        at = np.minimum(np.searchsorted(red, want), max(len(red) - 1, 0))  # red is sorted: every wanted value must be in it
        if len(red) == 0 or not np.array_equal(red[at], want) or not np.all(red[1:] > red[:-1]):
            raise RuntimeError("PARITY FAILED on the reduced host path (b200sk_run_reduced)")
        m = len(red)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            red = ctx.run_reduced(p, hb, ho, scale=args.scale, unique=True, copy=False)
        dt2 = time.perf_counter() - t0
        t = torch.tensor([dt2], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt2 = float(t.item())
        out["reduced"] = {"value": int(tot.item()) * args.e2e_steps / dt2, "unit": "bases/s", "scale": args.scale,
                          "api": "b200sk_run_reduced (keep h <= MaxUint64/scale, sort, unique on the device)",
                          "h2d_bytes_per_step": nb + (n + 1) * 8, "d2h_bytes_per_step": m * 8,
                          "ms_per_step": dt2 / args.e2e_steps * 1e3, "distinct_kept": m}
    L.b200sk_free_pinned(hb_ptr)
    L.b200sk_free_pinned(ho_ptr)
    if prev_affinity is not None:
        shard.restore_host_binding(prev_affinity)
    return out


def run_secondary(args, ctx, cabi, synth, oracle, rank, world, dev, dist, torch, np, peak, parity, cores, barrier, allmax,
                  allsum_i64):
    """BASELINE.json configs 2, 4 and 5, each sharded over the ranks, each with a sampled oracle check first."""
    L = cabi.lib()
    out = []
    steps = max(3, args.steps // 2)

    def one(name, label, p, omode, okw, bases, off, nb, n, frames=None, pos_bytes=4, note=None):
        exact = 1 if p.mode in (cabi.MODE_KMER, cabi.MODE_NTHASH, cabi.MODE_PROTEIN, cabi.MODE_SIMHASH) else 0
        cap = int(L.b200sk_output_bound(ctypes.byref(p), nb, n, exact))
        val = torch.empty(cap, dtype=torch.int64, device=dev)
        pos = torch.empty(cap, dtype=torch.int32, device=dev) if p.want_pos else None
        ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
        st = torch.empty(n, dtype=torch.int32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        plist = [p]
        if frames:
            plist = []
            for fr in frames:
                q = cabi.Params.from_buffer_copy(p)
                q.frame = fr
                plist.append(q)
        totals, mis, checked, ties = 0, 0, 0, 0
        for q in plist:  # parity first
            rc, tot = ctx.run_device(q, bases, off, nb, val, pos, ooff, st)
            if rc != 0:
                raise RuntimeError("%s: capacity" % name)
            totals += tot
            kw = dict(okw)
            if frames:
                kw["frame"] = int(q.frame)
            pr = parity_sample(torch, np, oracle, label, omode, kw, bases, off, val, pos, ooff, st,
                               max(1, args.parity_reads // world // len(plist)), 7 + rank, cores)
            mis += pr["mismatches"]; checked += pr["reads_checked"]; ties += pr["first_window_ties"]
        pr = {"config": label, "reads_checked": allsum_i64(checked), "mismatches": allsum_i64(mis),
              "first_window_ties": allsum_i64(ties)}
        if note:
            pr["note"] = note
        parity.append(pr)
        if pr["mismatches"]:
            raise RuntimeError("PARITY FAILED on %s: %r" % (name, pr))

        fused = None
        if frames:
            # the six frames through ONE call (b200sk_enqueue_device_frames: reads fetched and decoded once, walked six
            # times by one kernel), every frame into arrays of its own; checked against the per-frame launches above
            fused = dict(val=[torch.empty(cap, dtype=torch.int64, device=dev) for _ in frames],
                         off=[torch.empty(n + 1, dtype=torch.int64, device=dev) for _ in frames],
                         st=[torch.empty(n, dtype=torch.int32, device=dev) for _ in frames])
            ctx.enqueue_device_frames(p, bases, off, nb, fused["val"], fused["off"], fused["st"], flags)
            for fi, q in enumerate(plist):
                rc, tot = ctx.run_device(q, bases, off, nb, val, pos, ooff, st)
                same = (torch.equal(fused["off"][fi], ooff) and torch.equal(fused["st"][fi], st)
                        and torch.equal(fused["val"][fi][:tot], val[:tot]))
                if not same:
                    raise RuntimeError("PARITY FAILED on %s: fused six-frame call differs in frame %d" % (name, q.frame))

        def step():
            if fused:
                ctx.enqueue_device_frames(p, bases, off, nb, fused["val"], fused["off"], fused["st"], flags)
                return
            for q in plist:
                ctx.enqueue_device(q, bases, off, nb, val, pos, ooff, st, flags)

        l0 = ctx.kernel_launches()
        step()
        launches_per_step = ctx.kernel_launches() - l0
        ctx.timing_enable(True)
        ms = allmax(timed(torch, step, steps, 3, barrier))
        ksum, kn = ctx.timing_collect()
        ctx.timing_enable(False)
        kern_ms = allmax(ksum / max(kn, 1) * (1 if fused else len(plist)))
        tb, tr, to = allsum_i64(nb), allsum_i64(n), allsum_i64(totals)
        alg = tb + 8 * tr + (8 + (pos_bytes if p.want_pos else 0)) * to + 8 * tr
        out.append({"config": name, "value": tb / (ms * 1e-3), "unit": "bases/s", "ms_per_step": ms, "reads": tr,
                    "bases": tb, "elements": to, "launches_per_step": launches_per_step,
                    "roofline": {"bound": "hbm", "achieved": alg / world / (kern_ms * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s per GPU", "frac": alg / world / (kern_ms * 1e-3) / 1e9 / peak,
                                 "kernel_ms": kern_ms, "algorithmic_bytes": alg, "bytes_per_base": alg / tb}})
        del val, pos, ooff, st, fused

    # C2: canonical ntHash k=21, 10 M x 150 bp (values only: Index() of a dense mode is the running position)
    b = shard_bounds(args.c2_reads, world)
    n = b[rank + 1] - b[rank]
    bases, off = gen_uniform_shard(torch, b[rank], b[rank + 1], READ_LEN, 42, dev)
    one("C2 ntHash k=21 canonical, %d x 150 bp" % args.c2_reads, "C2 ntHash k=21",
        cabi.make_params(cabi.MODE_NTHASH, K, max_read_len=READ_LEN, want_pos=False), oracle.MODE_NTHASH, dict(k=K),
        bases, off, n * READ_LEN, n)
    del bases, off
    # C5: ProteinIterator k=11 over the six frames of 150-bp reads (one unit = six launches over the same reads)
    b = shard_bounds(args.c5_reads, world)
    n = b[rank + 1] - b[rank]
    bases, off = gen_uniform_shard(torch, b[rank], b[rank + 1], READ_LEN, 45, dev)
    one("C5 ProteinIterator k=11, six frames in one fused launch, %d x 150 bp" % args.c5_reads, "C5 protein k=11 x 6 frames",
        cabi.make_params(cabi.MODE_PROTEIN, 11, max_read_len=READ_LEN, want_pos=False, frame=1), oracle.MODE_PROTEIN,
        dict(k=11), bases, off, n * READ_LEN, n, frames=(1, 2, 3, -1, -2, -3),
        note="WYHASH_UNPINNED: GPU == oracle bit-exact; the oracle's wyhash is a restatement no reference vector pins")
    del bases, off
    torch.cuda.empty_cache()
    # C4: closed syncmers k=21 s=11 over ONT-like reads (lognormal lengths, mean 10 kb), sharded by cumulative bases
    lens = synth.ont_like_lengths(args.c4_reads, 44)
    ro = np.zeros(args.c4_reads + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    cut = cabi.shard_by_bases(ro, world)
    a0, a1 = int(cut[rank]), int(cut[rank + 1])
    n = a1 - a0
    nb = int(ro[a1] - ro[a0])
    g = torch.Generator(device=dev)
    g.manual_seed(4400 + rank)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = torch.zeros(nb + 64, dtype=torch.uint8, device=dev)
    stepb = 1 << 28
    for s0 in range(0, nb, stepb):
        e0 = min(nb, s0 + stepb)
        bases[s0:e0] = lut[torch.randint(0, 4, (e0 - s0,), generator=g, device=dev, dtype=torch.int64)]
    off = torch.from_numpy((ro[a0:a1 + 1] - ro[a0]).astype(np.int64)).to(dev)
    one("C4 closed syncmer k=21 s=11, %d ONT-like reads (mean 10 kb)" % args.c4_reads, "C4 syncmer k=21 s=11 ONT",
        cabi.make_params(cabi.MODE_SYNCMER, K, s=S, max_read_len=int(lens.max())), oracle.MODE_SYNCMER, dict(k=K, s=S),
        bases, off, nb, n)
    del bases, off
    torch.cuda.empty_cache()
    return out


def cpu_baseline(args):
    import oracle
    from bio_b200 import synth
    cores = os.cpu_count() or 1
    n = args.cpu_reads
    bases, off = synth.uniform_reads(n, READ_LEN, SEED)
    oracle.run_batch(bases[: 1000 * READ_LEN], off[:1001], oracle.MODE_MINIMIZER, k=K, w=W, threads=cores,
                     want_output=False)
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.run_batch(bases, off, oracle.MODE_MINIMIZER, k=K, w=W, threads=cores, want_output=False)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = time.perf_counter() - t0
    return {"value": n * READ_LEN * reps / dt, "unit": "bases/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x ({n} x {READ_LEN} bp uniform ACGT reads), oracle NextMinimizer restatement, "
                      f"{cores} threads over contiguous read shards",
            "note": "C restatement of the Go loops (no Go toolchain here); allocates per read where Go pools"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # launched without torchrun: re-launch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
