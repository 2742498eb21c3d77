#!/usr/bin/env python
"""bench.py -- bases/sec sketched (k=21, w=11 minimizer) on synthetic 150 bp reads, 1..8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the sketching hot path (sketches.NewMinimizerSketch + NextMinimizer/Index over
every read, reference: sketches/sketch.go:85,205) over one batch of synthetic reads, through the
C ABI of libb200sketch.so.  Records shard over ranks with no data-path collective ("weak": every
GPU holds one C3-sized shard).

JSON line (rank 0):
  value      whole-job bases/s with the batch already resident in HBM (device entry point)
  e2e        same metric through the host entry point b200sk_run: pinned HOST buffers in, H2D and D2H
             copies inside the timed region
  roofline   algorithmic bytes of the sketching kernel / its CUDA-event duration vs measured HBM copy peak
  cpu_baseline  the oracle (C restatement of the Go loops -- Go is absent here) on the host cores,
             bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W, READ_LEN = 21, 11, 150
SEED = 43
METRIC = "bases/sec sketched (k=21,w=11 minimizer)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU per step (C3: 100M x 150 bp)")
    ap.add_argument("--e2e-reads", type=int, default=0, help="reads per GPU per e2e step (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=2_000_000, help="reads in the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
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


def algorithmic_bytes(n_reads, n_bases, n_out):
    # each base read once (1 B ASCII), 8 B read offset in; 8 B value + 4 B position per
    # emitted element and 8 B output offset per read out.
    return n_bases + 8 * n_reads + 12 * n_out + 8 * n_reads


# ---------------------------------------------------------------- reference arm (CPU)
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  The reference is pure Go and neither a Go
    toolchain nor its four un-vendored arithmetic modules exist in this image, so this arm times the
    oracle: the line-by-line C restatement of sketches/sketch.go:205-309 (+ ntHash), one thread per host
    core over contiguous read shards -- kind "port"."""
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
    sample = f"{n} x {READ_LEN} bp uniform ACGT reads per step (seed {SEED}), same k/w; full workload is {args.reads} reads/GPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"C3 minimizer k={K} w={W}, {READ_LEN} bp reads, CPU sample", "reads_per_step": n},
        "cpu_baseline": {"value": v, "unit": "bases/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the Go algorithm (oracle/), not the Go binary: no Go toolchain "
                                 "in the image; published Go figure 18.3 Mbases/s/core (k=31,w=15, Ryzen 2700X)"},
        "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import ctypes
    import numpy as np
    import torch
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

    n = args.reads
    nb = n * READ_LEN
    bases, off = synth.device_uniform_reads(n, READ_LEN, SEED + rank, dev)
    p = cabi.make_params(cabi.MODE_MINIMIZER, K, w=W, max_read_len=READ_LEN)
    cap = int(cabi.lib().b200sk_output_bound(ctypes.byref(p), nb, n, 0))
    val = torch.empty(cap, dtype=torch.int64, device=dev)
    pos = torch.empty(cap, dtype=torch.int32, device=dev)
    ooff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)

    rc, n_out = ctx.run_device(p, bases, off, nb, val, pos, ooff, status)  # also sizes the scratch
    if rc != 0:
        raise RuntimeError("capacity estimate too small: need %d" % n_out)

    def step():
        ctx.enqueue_device(p, bases, off, nb, val, pos, ooff, status, flags)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.timing_enable(True)
    l0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - l0
    kern_ms_sum, kern_n = ctx.timing_collect()
    ctx.timing_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    if int(flags.item()) != 0:
        raise RuntimeError("kernel flags %d" % int(flags.item()))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nb * world / (ms_step * 1e-3)

    # roofline of the dominant (sketching) kernel on this rank
    peak, peak_src = measured_peak()
    kern_ms = kern_ms_sum / max(kern_n, 1)
    alg = algorithmic_bytes(n, nb, n_out)
    achieved = alg / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    if traffic and int(traffic.get("reads_per_launch", -1)) != n:
        traffic = None  # the committed capture is for another launch size
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "kernel": "k_sparse_warp<MINIMIZER,W=11>", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg,
                "bytes_per_base": alg / nb, "peak_source": peak_src,
                "kernel_share_of_step": kern_ms / ms_step}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")

    # optional: NCCL gather of the per-GPU uint64 hash arrays to rank 0 (not part of `value`)
    gather = None
    if dist is not None and not args.no_gather:
        from bio_b200 import shard
        part = val[: max(n_out // world, 1)]
        gather = shard.timed_gather(part, dist, dev)

    # end-to-end through the host entry point
    e2e = None
    if not args.no_e2e:
        del val, pos, ooff, status
        torch.cuda.empty_cache()
        e2e = run_e2e(args, ctx, cabi, synth, bases, rank, world, dev, dist, torch, np)
    del bases
    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_baseline(args)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C3: NewMinimizerSketch k={K} w={W} over {n} x {READ_LEN} bp uniform-ACGT reads per GPU",
                       "reads_per_gpu": n, "read_len": READ_LEN, "k": K, "w": W, "minimizers_per_read": n_out / n,
                       "parallelism": f"records sharded over {world} GPU(s), no data-path collective",
                       "l2": "inputs (15 GB) and outputs (27 GB) per step far exceed the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if gather:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, ctx, cabi, synth, d_bases, rank, world, dev, dist, torch, np):
    """Same metric through b200sk_run: inputs in pinned HOST memory, every step copies them to the device,
    sketches, and copies values, positions, offsets and statuses back to pinned host memory."""
    import ctypes
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
        n = int(min(args.reads, 0.45 * avail / world / per_read))
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
    res = None

    def step():
        return ctx.run(p, hb, ho, copy=False)

    res = step()  # warm-up: allocates the library's device + pinned output buffers
    res = step()
    n_out = res["total"]
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
    chk = int(res["val"][:1000].sum()) if n_out else 0
    out = {"value": nb * world * args.e2e_steps / dt, "unit": "bases/s",
           "h2d_bytes_per_step": nb + (n + 1) * 8,
           "d2h_bytes_per_step": n_out * 9 + (n + 1) * 8 + n * 4, "pos_width": 1,
           "reads_per_gpu_per_step": n, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
           "api": "b200sk_run (host pointers, pinned)", "checksum_first_1000": chk, "host_binding": binding}
    L.b200sk_free_pinned(hb_ptr)
    L.b200sk_free_pinned(ho_ptr)
    if prev_affinity is not None:
        shard.restore_host_binding(prev_affinity)
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
                      f"{cores} threads over contiguous read shards"}


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
