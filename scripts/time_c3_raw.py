"""A/B timing of a C3-geometry launch against ANY build of the library (dev helper; binds only the four entry points it
uses, so that an older libb200sketch.so can be loaded): B200SK_LIB_PATH=... python scripts/time_c3_raw.py [reads] [mode]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bio_b200 import _cabi as cabi, synth
path = os.environ.get("B200SK_LIB_PATH") or cabi.LIB_PATH
L = C.CDLL(path)
vp = C.c_void_p
L.b200sk_create.argtypes = [C.POINTER(vp), C.c_int]
L.b200sk_output_bound.restype = C.c_uint64
L.b200sk_output_bound.argtypes = [C.POINTER(cabi.Params), C.c_uint64, C.c_uint64, C.c_int]
L.b200sk_run_device.argtypes = [vp, C.POINTER(cabi.Params), vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, C.c_uint64, vp, C.POINTER(C.c_uint64)]
L.b200sk_enqueue_device.argtypes = [vp, C.POINTER(cabi.Params), vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, C.c_uint64, vp, vp]
dev = torch.device("cuda:0")
h = vp()
assert L.b200sk_create(C.byref(h), 0) == 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "minimizer"
bases, off = synth.device_uniform_reads(n, 150, 43, dev)
nb = n * 150
p = cabi.make_params(cabi.MODE_MINIMIZER, 21, w=11, max_read_len=150) if mode == "minimizer" else cabi.make_params(cabi.MODE_SYNCMER, 21, s=11, max_read_len=150)
cap = int(L.b200sk_output_bound(C.byref(p), nb, n, 0))
val = torch.empty(cap, dtype=torch.int64, device=dev); pos = torch.empty(cap, dtype=torch.int32, device=dev)
ooff = torch.empty(n + 1, dtype=torch.int64, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
flags = torch.zeros(1, dtype=torch.int32, device=dev)
tot = C.c_uint64(0)
s = torch.cuda.current_stream(dev).cuda_stream
assert L.b200sk_run_device(h, C.byref(p), bases.data_ptr(), off.data_ptr(), n, nb, val.data_ptr(), pos.data_ptr(), ooff.data_ptr(), st.data_ptr(), cap, s, C.byref(tot)) == 0
def step():
    assert L.b200sk_enqueue_device(h, C.byref(p), bases.data_ptr(), off.data_ptr(), n, nb, val.data_ptr(), pos.data_ptr(), ooff.data_ptr(), st.data_ptr(), cap, s, flags.data_ptr()) == 0
for _ in range(3): step()
torch.cuda.synchronize()
best = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 10)
print(json.dumps({"lib": os.path.basename(path), "mode": mode, "reads": n, "ms": [round(x, 4) for x in best], "Gbases_per_s": round(nb / min(best) / 1e6, 1)}), flush=True)
