"""ctypes binding of bio_b200/lib/libb200sketch.so (C ABI: include/b200sketch.h).

This is the only way Python reaches the sketching path: there is no CPU
fallback.  If the library is missing it is built with nvcc (in-tree); if that
fails, or no CUDA device is present at call time, the error is raised to the
caller.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SK_LIB_PATH") or os.path.join(_HERE, "lib", "libb200sketch.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
ERR_INVALID_K = -1
ERR_SHORT_SEQ = -2
ERR_INVALID_W = -3
ERR_INVALID_S = -4
ERR_ILLEGAL_BASE = -5
ERR_K_OVERFLOW = -6
ERR_INVALID_FRAME = -7
ERR_CODON_TABLE = -8
ERR_TRANSLATE_SHORT = -9
ERR_INVALID_CODON = -10
ERR_INVALID_M = -11
ERR_INVALID_SCALE = -12
ERR_K_TOO_LARGE = -13
ERR_CUDA = -100
ERR_NO_DEVICE = -101
ERR_UNSUPPORTED = -102
ERR_CAPACITY = -103
ERR_BAD_ARG = -104
ERR_NOMEM = -105

MODE_KMER, MODE_NTHASH, MODE_MINIMIZER, MODE_SYNCMER, MODE_PROTEIN, MODE_PROTEIN_MINIMIZER, MODE_SIMHASH = range(7)
ALPHABET_DNA_REDUNDANT, ALPHABET_DNA, ALPHABET_RNA_REDUNDANT, ALPHABET_RNA, ALPHABET_UNLIMIT, ALPHABET_PROTEIN = range(6)

# every symbol include/b200sketch.h declares (tests check the .so exports all of them)
EXPORTS = (
    "b200sk_create", "b200sk_destroy", "b200sk_alloc_pinned", "b200sk_free_pinned",
    "b200sk_check_params", "b200sk_output_bound", "b200sk_run", "b200sk_run_device",
    "b200sk_enqueue_device", "b200sk_strerror", "b200sk_last_error", "b200sk_kernel_launches",
    "b200sk_version", "b200sk_timing_enable", "b200sk_timing_collect",
    "b200sk_fastx_parse_device", "b200sk_run_fastx", "b200sk_copy_to_host",
    "b200sk_fxstream_open", "b200sk_fxstream_next", "b200sk_fxstream_rewind", "b200sk_fxstream_close",
    "b200sk_fxstream_kernel_launches", "b200sk_fxstream_last_error",
    "b200sk_gather_create", "b200sk_gather_open", "b200sk_gather_close", "b200sk_compact_segments",
    "b200sk_shard_by_bases", "b200sk_group_create", "b200sk_group_destroy", "b200sk_group_size", "b200sk_group_run",
    "b200sk_group_last_error", "b200sk_group_kernel_launches",
    "b200sk_scale_max_hash", "b200sk_reduce_device", "b200sk_run_reduced", "b200sk_enqueue_device_sharded",
    "b200sk_enqueue_device_frames", "b200sk_run_frames",
)
IPC_HANDLE_BYTES = 64
FXSTREAM_END = 1

FASTX_FASTA, FASTX_FASTQ = 1, 2
ERR_NOT_FASTX, ERR_BAD_FASTQ = -20, -21


class FastxInfo(C.Structure):
    _fields_ = [
        ("format", C.c_int32), ("status", C.c_int32), ("n_records", C.c_uint64), ("n_bases", C.c_uint64),
        ("n_lines", C.c_uint64), ("consumed", C.c_uint64), ("bad_record", C.c_uint64),
        ("max_read_len", C.c_uint32), ("reserved", C.c_uint32),
        ("d_bases", C.c_void_p), ("d_read_off", C.c_void_p), ("d_rec_off", C.c_void_p),
        ("d_qual_off", C.c_void_p), ("d_line_off", C.c_void_p),
        ("alphabet", C.c_int32), ("reserved2", C.c_int32), ("first_invalid", C.c_uint64), ("d_invalid", C.c_void_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("k", C.c_int32), ("w", C.c_int32), ("s", C.c_int32),
        ("canonical", C.c_int32), ("circular", C.c_int32), ("codon_table", C.c_int32),
        ("frame", C.c_int32), ("alphabet", C.c_int32), ("want_pos", C.c_int32),
        ("max_read_len", C.c_uint32), ("m", C.c_int32), ("scale", C.c_int32), ("pos_width", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


class ShardSpec(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("chunk_reads", C.c_uint32), ("epoch", C.c_uint32),
                ("n_reads_global", C.c_uint64), ("state", C.c_void_p * 8)]


def make_params(mode, k, w=0, s=0, canonical=True, circular=False, codon_table=1, frame=1,
                alphabet=ALPHABET_DNA_REDUNDANT, want_pos=True, max_read_len=0, m=0, scale=1, pos_width=0):
    p = Params()
    p.mode, p.k, p.w, p.s = mode, k, w, s
    p.canonical, p.circular = int(bool(canonical)), int(bool(circular))
    p.codon_table, p.frame, p.alphabet = codon_table, frame, alphabet
    p.want_pos, p.max_read_len = int(bool(want_pos)), int(max_read_len)
    p.m, p.scale, p.pos_width = m, scale, pos_width
    return p


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", CSRC], stdout=None if verbose else subprocess.DEVNULL)


def _needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")) and os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return False


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    alt = os.environ.get("B200SK_LIB_PATH")  # development only: an A/B build of the same library (scripts/ab_build.sh)
    if not alt and _needs_build():
        build()
    L = C.CDLL(alt or LIB_PATH)
    vp, u8p, u64p, u32p, i32p = (C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
    PP = C.POINTER(Params)
    L.b200sk_version.restype = C.c_int
    L.b200sk_create.restype = C.c_int
    L.b200sk_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.b200sk_destroy.restype = None
    L.b200sk_destroy.argtypes = [vp]
    L.b200sk_alloc_pinned.restype = C.c_void_p
    L.b200sk_alloc_pinned.argtypes = [C.c_size_t]
    L.b200sk_free_pinned.restype = None
    L.b200sk_free_pinned.argtypes = [vp]
    L.b200sk_check_params.restype = C.c_int
    L.b200sk_check_params.argtypes = [PP]
    L.b200sk_output_bound.restype = C.c_uint64
    L.b200sk_output_bound.argtypes = [PP, C.c_uint64, C.c_uint64, C.c_int]
    L.b200sk_run.restype = C.c_int
    L.b200sk_run.argtypes = [vp, PP, u8p, u64p, C.c_uint64,
                             C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                             C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.b200sk_run_device.restype = C.c_int
    L.b200sk_run_device.argtypes = [vp, PP, u8p, u64p, C.c_uint64, C.c_uint64, u64p, u32p, u64p, i32p,
                                    C.c_uint64, vp, C.POINTER(C.c_uint64)]
    L.b200sk_enqueue_device.restype = C.c_int
    L.b200sk_enqueue_device.argtypes = [vp, PP, u8p, u64p, C.c_uint64, C.c_uint64, u64p, u32p, u64p, i32p,
                                        C.c_uint64, vp, u32p]
    L.b200sk_strerror.restype = C.c_char_p
    L.b200sk_strerror.argtypes = [C.c_int]
    L.b200sk_last_error.restype = C.c_char_p
    L.b200sk_last_error.argtypes = [vp]
    L.b200sk_kernel_launches.restype = C.c_uint64
    L.b200sk_kernel_launches.argtypes = [vp]
    L.b200sk_timing_enable.restype = None
    L.b200sk_timing_enable.argtypes = [vp, C.c_int]
    L.b200sk_timing_collect.restype = C.c_int
    L.b200sk_timing_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.b200sk_fastx_parse_device.restype = C.c_int
    L.b200sk_fastx_parse_device.argtypes = [vp, u8p, C.c_uint64, C.c_int, C.c_int, vp, C.POINTER(FastxInfo)]
    L.b200sk_copy_to_host.restype = C.c_int
    L.b200sk_copy_to_host.argtypes = [vp, vp, vp, C.c_uint64]
    L.b200sk_run_fastx.restype = C.c_int
    L.b200sk_run_fastx.argtypes = [vp, PP, u8p, C.c_uint64, C.c_int, C.c_int, C.POINTER(FastxInfo),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.b200sk_fxstream_open.restype = C.c_int
    L.b200sk_fxstream_open.argtypes = [C.POINTER(C.c_void_p), C.c_int, PP, u8p, C.c_uint64, C.c_int, C.c_uint64]
    L.b200sk_fxstream_next.restype = C.c_int
    L.b200sk_fxstream_next.argtypes = [vp, C.POINTER(FastxInfo), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.b200sk_fxstream_rewind.restype = C.c_int
    L.b200sk_fxstream_rewind.argtypes = [vp, u8p, C.c_uint64, C.c_int]
    L.b200sk_fxstream_close.restype = None
    L.b200sk_fxstream_close.argtypes = [vp]
    L.b200sk_fxstream_kernel_launches.restype = C.c_uint64
    L.b200sk_fxstream_kernel_launches.argtypes = [vp]
    L.b200sk_fxstream_last_error.restype = C.c_char_p
    L.b200sk_fxstream_last_error.argtypes = [vp]
    L.b200sk_gather_create.restype = C.c_int
    L.b200sk_gather_create.argtypes = [vp, C.c_uint64, vp, C.POINTER(C.c_void_p)]
    L.b200sk_gather_open.restype = C.c_int
    L.b200sk_gather_open.argtypes = [vp, vp, C.POINTER(C.c_void_p)]
    L.b200sk_gather_close.restype = C.c_int
    L.b200sk_gather_close.argtypes = [vp, vp, C.c_int]
    L.b200sk_compact_segments.restype = C.c_int
    L.b200sk_compact_segments.argtypes = [vp, vp, u64p, u64p, C.c_int, vp]
    L.b200sk_shard_by_bases.restype = None
    L.b200sk_shard_by_bases.argtypes = [u64p, C.c_uint64, C.c_int, u64p]
    L.b200sk_group_create.restype = C.c_int
    L.b200sk_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    L.b200sk_group_destroy.restype = None
    L.b200sk_group_destroy.argtypes = [vp]
    L.b200sk_group_size.restype = C.c_int
    L.b200sk_group_size.argtypes = [vp]
    L.b200sk_group_run.restype = C.c_int
    L.b200sk_group_run.argtypes = [vp, PP, u8p, u64p, C.c_uint64,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.b200sk_group_last_error.restype = C.c_char_p
    L.b200sk_group_last_error.argtypes = [vp]
    L.b200sk_group_kernel_launches.restype = C.c_uint64
    L.b200sk_group_kernel_launches.argtypes = [vp]
    L.b200sk_scale_max_hash.restype = C.c_uint64
    L.b200sk_scale_max_hash.argtypes = [C.c_uint32]
    L.b200sk_enqueue_device_sharded.restype = C.c_int
    L.b200sk_enqueue_device_sharded.argtypes = [vp, PP, C.POINTER(ShardSpec), u8p, u64p, C.c_uint64, C.c_uint64, vp, vp, vp,
                                                vp, C.c_uint64, vp, u32p]
    L.b200sk_enqueue_device_frames.restype = C.c_int
    L.b200sk_enqueue_device_frames.argtypes = [vp, PP, u8p, u64p, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint64, vp, u32p]
    L.b200sk_run_frames.restype = C.c_int
    L.b200sk_run_frames.argtypes = [vp, PP, u8p, u64p, C.c_uint64, vp, vp, vp, vp]
    L.b200sk_run_reduced.restype = C.c_int
    L.b200sk_run_reduced.argtypes = [vp, PP, C.c_uint32, C.c_int, u8p, u64p, C.c_uint64, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint64)]
    L.b200sk_reduce_device.restype = C.c_int
    L.b200sk_reduce_device.argtypes = [vp, u64p, C.c_uint64, C.c_uint32, C.c_int, u64p, C.c_uint64,
                                       C.POINTER(C.c_uint64), vp]
    _lib = L
    return L


def strerror(code):
    return lib().b200sk_strerror(code).decode()


class SketchError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = strerror(code)
        if detail:
            msg += " (" + detail + ")"
        super().__init__(msg)


class Context:
    """One b200sk_ctx: a device, its streams and scratch buffers."""

    def __init__(self, device=0):
        L = lib()
        h = C.c_void_p()
        rc = L.b200sk_create(C.byref(h), device)
        if rc != 0:
            raise SketchError(rc)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().b200sk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _raise(self, rc):
        raise SketchError(rc, lib().b200sk_last_error(self._h).decode() if rc == ERR_CUDA else "")

    def timing_enable(self, on=True):
        lib().b200sk_timing_enable(self._h, int(on))

    def timing_collect(self):
        """(sum of main-kernel durations in ms, number of kernels) since the last collect."""
        s, n = C.c_double(0), C.c_uint64(0)
        rc = lib().b200sk_timing_collect(self._h, C.byref(s), C.byref(n))
        if rc != 0:
            self._raise(rc)
        return s.value, int(n.value)

    def kernel_launches(self):
        return int(lib().b200sk_kernel_launches(self._h))

    # ---- host entry point: numpy in, numpy views of library-owned pinned memory out
    def run(self, params, bases, read_off, copy=True):
        import numpy as np
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        ov, op, oo, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        total = C.c_uint64(0)
        rc = lib().b200sk_run(self._h, C.byref(params), bases.ctypes.data, read_off.ctypes.data, n,
                              C.byref(ov), C.byref(op), C.byref(oo), C.byref(st), C.byref(total))
        if rc != 0:
            self._raise(rc)
        t = int(total.value)

        def view(ptr, count, dt):
            if not ptr.value or count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dt).itemsize,))
            a = a.view(dt)
            return a.copy() if copy else a

        pdt = {1: np.uint8, 2: np.uint16}.get(int(params.pos_width), np.uint32)
        return dict(val=view(ov, t, np.uint64), pos=view(op, t, pdt) if params.want_pos else None,
                    off=view(oo, n + 1, np.uint64), status=view(st, n, np.int32), total=t)

    # ---- device entry points: torch CUDA tensors (plumbing only: pointers + the current stream)
    def run_device(self, params, d_bases, d_off, n_bases, out_val, out_pos, out_off, status, stream=None):
        import torch
        n = d_off.numel() - 1
        st = torch.cuda.current_stream(d_bases.device).cuda_stream if stream is None else stream
        total = C.c_uint64(0)
        rc = lib().b200sk_run_device(
            self._h, C.byref(params), d_bases.data_ptr(), d_off.data_ptr(), n, n_bases,
            out_val.data_ptr() if out_val is not None else None,
            out_pos.data_ptr() if out_pos is not None else None,
            out_off.data_ptr(), status.data_ptr() if status is not None else None,
            out_val.numel() if out_val is not None else 0, st, C.byref(total))
        if rc not in (0, ERR_CAPACITY):
            self._raise(rc)
        return rc, int(total.value)

    def enqueue_device(self, params, d_bases, d_off, n_bases, out_val, out_pos, out_off, status, flags,
                       stream=None):
        import torch
        n = d_off.numel() - 1
        st = torch.cuda.current_stream(d_bases.device).cuda_stream if stream is None else stream
        rc = lib().b200sk_enqueue_device(
            self._h, C.byref(params), d_bases.data_ptr(), d_off.data_ptr(), n, n_bases,
            out_val.data_ptr() if out_val is not None else None,
            out_pos.data_ptr() if out_pos is not None else None,
            out_off.data_ptr(), status.data_ptr() if status is not None else None,
            out_val.numel() if out_val is not None else 0, st,
            flags.data_ptr() if flags is not None else None)
        if rc != 0:
            self._raise(rc)

    # ---- multi-GPU gather buffer (CUDA IPC): raw device addresses, the caller wraps them as it likes
    def gather_create(self, capacity_elems):
        """Root: allocate the gather buffer.  Returns (64-byte handle, device address)."""
        h = (C.c_uint8 * IPC_HANDLE_BYTES)()
        d = C.c_void_p()
        rc = lib().b200sk_gather_create(self._h, int(capacity_elems), h, C.byref(d))
        if rc != 0:
            self._raise(rc)
        return bytes(h), int(d.value)

    def gather_open(self, handle):
        """Other ranks: map the root's buffer into this process.  Returns the device address."""
        h = (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle)
        d = C.c_void_p()
        rc = lib().b200sk_gather_open(self._h, h, C.byref(d))
        if rc != 0:
            self._raise(rc)
        return int(d.value)

    def gather_close(self, addr, is_owner):
        rc = lib().b200sk_gather_close(self._h, C.c_void_p(addr), int(bool(is_owner)))
        if rc != 0:
            self._raise(rc)

    def compact_segments(self, addr, seg_base, seg_count, stream=None):
        import numpy as np
        b = np.ascontiguousarray(seg_base, dtype=np.uint64)
        c = np.ascontiguousarray(seg_count, dtype=np.uint64)
        rc = lib().b200sk_compact_segments(self._h, C.c_void_p(addr), b.ctypes.data, c.ctypes.data, len(b),
                                           C.c_void_p(stream or 0))
        if rc != 0:
            self._raise(rc)

    def enqueue_device_raw(self, params, d_bases, d_off, n_bases, out_val_addr, capacity, out_pos, out_off, status,
                           flags, stream=None):
        """enqueue_device with out_val given as a raw device address (a peer-mapped segment of the gather buffer)."""
        import torch
        n = d_off.numel() - 1
        st = torch.cuda.current_stream(d_bases.device).cuda_stream if stream is None else stream
        rc = lib().b200sk_enqueue_device(
            self._h, C.byref(params), d_bases.data_ptr(), d_off.data_ptr(), n, n_bases, C.c_void_p(out_val_addr),
            out_pos.data_ptr() if out_pos is not None else None, out_off.data_ptr(),
            status.data_ptr() if status is not None else None, int(capacity), st,
            flags.data_ptr() if flags is not None else None)
        if rc != 0:
            self._raise(rc)

    def enqueue_device_sharded(self, params, spec, d_bases, d_off, n_bases, val_addr, pos_addr, off_addr, status_addr,
                               capacity, flags, stream=None):
        """One rank's part of a batch sharded over several GPUs: the output arrays are raw device addresses of the
        ROOT's buffers (peer-mapped here), indexed globally (include/b200sketch.h: b200sk_enqueue_device_sharded)."""
        import torch
        n = d_off.numel() - 1
        st = torch.cuda.current_stream(d_bases.device).cuda_stream if stream is None else stream
        rc = lib().b200sk_enqueue_device_sharded(
            self._h, C.byref(params), C.byref(spec), d_bases.data_ptr(), d_off.data_ptr(), n, n_bases,
            C.c_void_p(val_addr), C.c_void_p(pos_addr) if pos_addr else None, C.c_void_p(off_addr),
            C.c_void_p(status_addr) if status_addr else None, int(capacity), st,
            flags.data_ptr() if flags is not None else None)
        if rc != 0:
            self._raise(rc)

    def enqueue_device_frames(self, params, d_bases, d_off, n_bases, out_vals, out_offs, statuses, flags, stream=None):
        """ProteinIterator over all six frames (1, 2, 3, -1, -2, -3) of every read: six value / offset / status
        tensors in, one call (include/b200sketch.h: b200sk_enqueue_device_frames)."""
        import torch
        n = d_off.numel() - 1
        st = torch.cuda.current_stream(d_bases.device).cuda_stream if stream is None else stream
        vals = (C.c_void_p * 6)(*[t.data_ptr() for t in out_vals])
        offs = (C.c_void_p * 6)(*[t.data_ptr() for t in out_offs])
        sts = (C.c_void_p * 6)(*[t.data_ptr() for t in statuses]) if statuses is not None else None
        rc = lib().b200sk_enqueue_device_frames(
            self._h, C.byref(params), d_bases.data_ptr(), d_off.data_ptr(), n, n_bases, vals, offs, sts,
            min(t.numel() for t in out_vals), st, flags.data_ptr() if flags is not None else None)
        if rc != 0:
            self._raise(rc)

    def run_frames(self, params, bases, read_off, copy=True):
        """Host entry point of the six-frame ProteinIterator call: a list of six dicts (val, off, status, total) for
        frames 1, 2, 3, -1, -2, -3 (include/b200sketch.h: b200sk_run_frames)."""
        import numpy as np
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        ov, oo, os_ = (C.c_void_p * 6)(), (C.c_void_p * 6)(), (C.c_void_p * 6)()
        tot = (C.c_uint64 * 6)()
        rc = lib().b200sk_run_frames(self._h, C.byref(params), bases.ctypes.data, read_off.ctypes.data, n, ov, oo, os_, tot)
        if rc != 0:
            self._raise(rc)

        def view(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dt).itemsize,)).view(dt)
            return a.copy() if copy else a
        return [dict(val=view(ov[i], int(tot[i]), np.uint64), pos=None, off=view(oo[i], n + 1, np.uint64),
                     status=view(os_[i], n, np.int32), total=int(tot[i])) for i in range(6)]

    def run_reduced(self, params, bases, read_off, scale=1, unique=True, copy=True):
        """Host entry point with the reduction inside: returns the sorted (distinct) values <= MaxUint64/scale."""
        import numpy as np
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        ov = C.c_void_p()
        total = C.c_uint64(0)
        rc = lib().b200sk_run_reduced(self._h, C.byref(params), int(scale), int(bool(unique)), bases.ctypes.data,
                                      read_off.ctypes.data, len(read_off) - 1, C.byref(ov), C.byref(total))
        if rc != 0:
            self._raise(rc)
        t = int(total.value)
        if not ov.value or t == 0:
            return np.zeros(0, dtype=np.uint64)
        a = np.ctypeslib.as_array(C.cast(ov, C.POINTER(C.c_uint64)), shape=(t,))
        return a.copy() if copy else a

    # ---- downstream reduction: FracMinHash filter + sort + unique on resident arrays
    def reduce_device(self, d_val, n, d_out, scale=1, unique=True, stream=None):
        """d_val[:n] (torch int64 CUDA tensor, consumed as scratch) -> d_out: values <= MaxUint64/scale, sorted, distinct.
        Returns (rc, number of elements produced or needed)."""
        import torch
        st = torch.cuda.current_stream(d_val.device).cuda_stream if stream is None else stream
        total = C.c_uint64(0)
        rc = lib().b200sk_reduce_device(self._h, d_val.data_ptr(), int(n), int(scale), int(bool(unique)),
                                        d_out.data_ptr(), d_out.numel(), C.byref(total), st)
        if rc not in (0, ERR_CAPACITY):
            self._raise(rc)
        return rc, int(total.value)

    # ---- record feeder (seqio/fastx.Reader as a batch operation)
    def fastx_parse_device(self, d_text, n_bytes=None, fmt=0, final=True, stream=None):
        """Split a chunk of FASTA/FASTQ text resident in HBM (torch uint8 CUDA tensor, 16-byte aligned, padded
        to a multiple of 16 bytes) into records.  Returns the FastxInfo (device pointers are library-owned)."""
        import torch
        n = d_text.numel() if n_bytes is None else n_bytes
        st = torch.cuda.current_stream(d_text.device).cuda_stream if stream is None else stream
        info = FastxInfo()
        rc = lib().b200sk_fastx_parse_device(self._h, d_text.data_ptr(), n, fmt, int(bool(final)), st, C.byref(info))
        if rc != 0:
            e = SketchError(rc, lib().b200sk_last_error(self._h).decode() if rc == ERR_CUDA else "")
            e.info = info
            raise e
        return info

    def fastx_fetch(self, info):
        """Copy a parse result to the host (tests): dict of numpy arrays."""
        import numpy as np

        def grab(ptr, count, dt):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dt)
            out = np.empty(count, dtype=dt)
            rc = lib().b200sk_copy_to_host(self._h, out.ctypes.data, ptr, out.nbytes)
            if rc != 0:
                self._raise(rc)
            return out

        n = int(info.n_records)
        return dict(format=int(info.format), n_records=n, consumed=int(info.consumed),
                    max_read_len=int(info.max_read_len),
                    bases=grab(info.d_bases, int(info.n_bases), np.uint8),
                    read_off=grab(info.d_read_off, n + 1, np.uint64),
                    rec_off=grab(info.d_rec_off, n + 1, np.uint64),
                    qual_off=grab(info.d_qual_off, n, np.uint64),
                    line_off=grab(info.d_line_off, int(info.n_lines) + 1, np.uint64),
                    alphabet=int(info.alphabet), first_invalid=int(info.first_invalid),
                    invalid=grab(info.d_invalid, n, np.uint8))

    def run_fastx(self, params, text, fmt=0, final=True, copy=True):
        """Host entry point: FASTA/FASTQ text (bytes / uint8 array) in, the sketches of its records out."""
        import numpy as np
        text = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) \
            else np.ascontiguousarray(text, dtype=np.uint8)
        info = FastxInfo()
        ov, op, oo, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        total = C.c_uint64(0)
        rc = lib().b200sk_run_fastx(self._h, C.byref(params), text.ctypes.data if len(text) else None, len(text), fmt,
                                    int(bool(final)), C.byref(info), C.byref(ov), C.byref(op), C.byref(oo),
                                    C.byref(st), C.byref(total))
        if rc != 0:
            self._raise(rc)
        t, n = int(total.value), int(info.n_records)

        def view(ptr, count, dt):
            if not ptr.value or count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dt).itemsize,))
            a = a.view(dt)
            return a.copy() if copy else a

        pdt = {1: np.uint8, 2: np.uint16}.get(int(params.pos_width), np.uint32)
        return dict(val=view(ov, t, np.uint64), pos=view(op, t, pdt) if params.want_pos else None,
                    off=view(oo, n + 1, np.uint64), status=view(st, n, np.int32), total=t, info=info)


def shard_by_bases(read_off, n_shards):
    """cut[n_shards + 1]: shard d owns reads [cut[d], cut[d+1]), balanced by cumulative bases (host logic, no device)."""
    import numpy as np
    ro = np.ascontiguousarray(read_off, dtype=np.uint64)
    cut = np.zeros(n_shards + 1, dtype=np.uint64)
    lib().b200sk_shard_by_bases(ro.ctypes.data, len(ro) - 1, n_shards, cut.ctypes.data)
    return cut


class Group:
    """b200sk_group: one process driving several devices (a context, stream and worker thread per device)."""

    def __init__(self, devices):
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = lib().b200sk_group_create(C.byref(h), arr, len(devices))
        if rc != 0:
            raise SketchError(rc)
        self._h = h

    def size(self):
        return int(lib().b200sk_group_size(self._h))

    def kernel_launches(self):
        return int(lib().b200sk_group_kernel_launches(self._h))

    def run(self, params, bases, read_off, copy=True):
        import numpy as np
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        ov, op, oo, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        total = C.c_uint64(0)
        rc = lib().b200sk_group_run(self._h, C.byref(params), bases.ctypes.data, read_off.ctypes.data, n,
                                    C.byref(ov), C.byref(op), C.byref(oo), C.byref(st), C.byref(total))
        if rc != 0:
            raise SketchError(rc, lib().b200sk_group_last_error(self._h).decode() if rc == ERR_CUDA else "")
        t = int(total.value)

        def view(ptr, count, dt):
            if not ptr.value or count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dt).itemsize,))
            a = a.view(dt)
            return a.copy() if copy else a

        pdt = {1: np.uint8, 2: np.uint16}.get(int(params.pos_width), np.uint32)
        return dict(val=view(ov, t, np.uint64), pos=view(op, t, pdt) if params.want_pos else None,
                    off=view(oo, n + 1, np.uint64), status=view(st, n, np.int32), total=t)

    def close(self):
        if getattr(self, "_h", None):
            lib().b200sk_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastxStream:
    """b200sk_fxstream: a whole FASTA/FASTQ text in host memory, sketched chunk by chunk with the copy and parse
    of the next chunk overlapping the sketching and copy back of this one.  Iterating yields one dict per chunk
    (as Context.run_fastx); with copy=False the arrays are views valid until the next chunk is asked for."""

    def __init__(self, params, text, device=0, fmt=0, chunk_bytes=0, copy=True):
        self._h = None
        self.params, self.copy = params, copy
        self._text = self._as_array(text)
        h = C.c_void_p()
        rc = lib().b200sk_fxstream_open(C.byref(h), device, C.byref(params), self._ptr(), len(self._text), fmt,
                                        chunk_bytes)
        if rc != 0:
            raise SketchError(rc)
        self._h = h

    @staticmethod
    def _as_array(text):
        import numpy as np
        return np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) \
            else np.ascontiguousarray(text, dtype=np.uint8)

    def _ptr(self):
        return self._text.ctypes.data if len(self._text) else None

    def rewind(self, text=None, fmt=0):
        old = self._text  # chunks of the old text may still be in flight until the call below returns
        if text is not None:
            self._text = self._as_array(text)
        rc = lib().b200sk_fxstream_rewind(self._h, self._ptr(), len(self._text), fmt)
        del old
        if rc != 0:
            raise SketchError(rc)

    def next(self):
        """The next chunk, or None after the last one."""
        import numpy as np
        info = FastxInfo()
        ov, op, oo, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        total = C.c_uint64(0)
        rc = lib().b200sk_fxstream_next(self._h, C.byref(info), C.byref(ov), C.byref(op), C.byref(oo), C.byref(st),
                                        C.byref(total))
        if rc == FXSTREAM_END:
            return None
        if rc != 0:
            e = SketchError(rc, lib().b200sk_fxstream_last_error(self._h).decode() if rc == -100 else "")
            e.info = info
            raise e
        t, n = int(total.value), int(info.n_records)

        def view(ptr, count, dt):
            if not ptr.value or count == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dt).itemsize,))
            a = a.view(dt)
            return a.copy() if self.copy else a

        pdt = {1: np.uint8, 2: np.uint16}.get(int(self.params.pos_width), np.uint32)
        return dict(val=view(ov, t, np.uint64), pos=view(op, t, pdt) if self.params.want_pos else None,
                    off=view(oo, n + 1, np.uint64), status=view(st, n, np.int32), total=t, info=info)

    def __iter__(self):
        while True:
            c = self.next()
            if c is None:
                return
            yield c

    @property
    def kernel_launches(self):
        return int(lib().b200sk_fxstream_kernel_launches(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200sk_fxstream_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
